// Input pipeline kernels (see input.cuh).
#include "input.cuh"

#include <algorithm>
#include <cmath>

#include "common.h"

namespace sam3b {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Pillow Resample.c

__device__ __forceinline__ int clip8(int v) { return min(max(v >> PRECISION_BITS, 0), 255); }

// horizontal pass: one thread per (row, output column), three channels
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, int h, int w, int out,
                                                       const int32_t* __restrict__ bounds, const int32_t* __restrict__ coeffs, int ks,
                                                       uint8_t* __restrict__ tmp) {
  const int64_t total = (int64_t)h * out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % out), y = (int)(i / out);
    const int lo = __ldg(bounds + 2 * X), n = __ldg(bounds + 2 * X + 1);
    const int32_t* k = coeffs + (int64_t)X * ks;
    const uint8_t* p = src + ((int64_t)y * w + lo) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int c = __ldg(k + t);
      a0 += p[3 * t] * c; a1 += p[3 * t + 1] * c; a2 += p[3 * t + 2] * c;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = (uint8_t)clip8(a0); o[1] = (uint8_t)clip8(a1); o[2] = (uint8_t)clip8(a2);
  }
}

// vertical pass + ToTensor + Normalize: one thread per output pixel, writes the three fp32 planes
__global__ void __launch_bounds__(256) resize_v_normalize_kernel(const uint8_t* __restrict__ tmp, int out_w, int out_h,
                                                                 const int32_t* __restrict__ bounds, const int32_t* __restrict__ coeffs,
                                                                 int ks, float mean, float std, float* __restrict__ dst) {
  const int64_t total = (int64_t)out_h * out_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % out_w), Y = (int)(i / out_w);
    const int lo = __ldg(bounds + 2 * Y), n = __ldg(bounds + 2 * Y + 1);
    const int32_t* k = coeffs + (int64_t)Y * ks;
    const uint8_t* p = tmp + ((int64_t)lo * out_w + X) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int c = __ldg(k + t);
      const uint8_t* q = p + (int64_t)t * out_w * 3;
      a0 += q[0] * c; a1 += q[1] * c; a2 += q[2] * c;
    }
    // ToTensor: float32(u8) / 255 ; Normalize: (t - mean) / std   (IEEE float32 division, as torch does)
    dst[i] = __fdiv_rn(__fdiv_rn((float)clip8(a0), 255.f) - mean, std);
    dst[total + i] = __fdiv_rn(__fdiv_rn((float)clip8(a1), 255.f) - mean, std);
    dst[2 * total + i] = __fdiv_rn(__fdiv_rn((float)clip8(a2), 255.f) - mean, std);
  }
}

// ATen nearest_idx with a float32 scale
__device__ __forceinline__ int nearest_idx(int dst, int in_size, int out_size) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  return min((int)floorf(__fmul_rn((float)dst, scale)), in_size - 1);
}

__global__ void __launch_bounds__(256) rle_masks_nearest_kernel(const uint32_t* __restrict__ cum, const int32_t* __restrict__ offs,
                                                                const int32_t* __restrict__ hw, int out, uint8_t* __restrict__ dst) {
  const int n = blockIdx.y;
  const int h = hw[2 * n], w = hw[2 * n + 1];
  const uint32_t* c = cum + offs[n];
  const int runs = offs[n + 1] - offs[n];
  const int64_t total = (int64_t)out * out;
  uint8_t* d = dst + (int64_t)n * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % out), Y = (int)(i / out);
    const uint32_t idx = (uint32_t)nearest_idx(X, w, out) * (uint32_t)h + (uint32_t)nearest_idx(Y, h, out);   // column-major
    // first run r with cum[r] > idx; odd runs are ones
    int lo = 0, hi = runs;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(c + mid) > idx) hi = mid; else lo = mid + 1;
    }
    d[i] = (uint8_t)((lo < runs) ? (lo & 1) : 0);
  }
}

// ---- polygon masks (pycocotools rleFrPoly): crossing points of the 5x up-sampled boundary, then parity fill ----------
// One thread per boundary point.  Point (edge e, step d) and its predecessor on the closed walk are recomputed from the
// edge's integer end points; double arithmetic uses explicit round-to-nearest mul / add / div (no FMA contraction), i.e.
// the operations a plain x86-64 build of maskApi.c executes.
struct PolyPoint { int u, v; };
__device__ __forceinline__ PolyPoint poly_point(int xs, int ys, int xe, int ye, int d) {
  const int dx = abs(xe - xs), dy = abs(ys - ye);
  const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
  if (flip) { int t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
  PolyPoint p;
  if (dx >= dy) {
    const double s = dx > 0 ? __ddiv_rn((double)(ye - ys), (double)dx) : 0.0;
    const int t = flip ? dx - d : d;
    p.u = t + xs;
    p.v = (int)__dadd_rn(__dadd_rn((double)ys, __dmul_rn(s, (double)t)), 0.5);
  } else {
    const double s = __ddiv_rn((double)(xe - xs), (double)dy);
    const int t = flip ? dy - d : d;
    p.v = t + ys;
    p.u = (int)__dadd_rn(__dadd_rn((double)xs, __dmul_rn(s, (double)t)), 0.5);
  }
  return p;
}

__global__ void __launch_bounds__(256) poly_crossings_kernel(const int32_t* __restrict__ edges, const int32_t* __restrict__ pt_start,
                                                             int n_edges, int64_t total_pts, int64_t* __restrict__ keys) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_pts; g += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_edges;                       // last edge e with pt_start[e] <= g
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(pt_start + mid) <= g) lo = mid; else hi = mid;
    }
    const int e = lo, d = (int)(g - __ldg(pt_start + e));
    const int32_t* E = edges + (int64_t)e * 8;       // xs, ys, xe, ye, list id, h, w, first edge of its polygon
    int64_t key = INT64_MAX;
    if (!(d == 0 && E[7] != 0)) {                   // the polygon's very first point has no predecessor
      const PolyPoint c = poly_point(E[0], E[1], E[2], E[3], d);
      PolyPoint q;
      if (d > 0) q = poly_point(E[0], E[1], E[2], E[3], d - 1);
      else {
        const int32_t* P = E - 8;                    // previous edge of the same polygon: its last point
        q = poly_point(P[0], P[1], P[2], P[3], max(abs(P[2] - P[0]), abs(P[1] - P[3])));
      }
      if (c.u != q.u) {
        const int h = E[5], w = E[6];
        double xd = (double)(c.u < q.u ? c.u : c.u - 1);
        xd = __dadd_rn(__ddiv_rn(__dadd_rn(xd, 0.5), 5.0), -0.5);
        if (floor(xd) == xd && xd >= 0.0 && xd <= (double)(w - 1)) {
          double yd = (double)(c.v < q.v ? c.v : q.v);
          yd = __dadd_rn(__ddiv_rn(__dadd_rn(yd, 0.5), 5.0), -0.5);
          yd = yd < 0.0 ? 0.0 : (yd > (double)h ? (double)h : yd);
          yd = ceil(yd);
          key = ((int64_t)E[4] << 32) | (int64_t)((int)xd * h + (int)yd);
        }
      }
    }
    keys[g] = key;
  }
}

// first index i with keys[i] >= k
__device__ __forceinline__ int64_t key_lower_bound(const int64_t* __restrict__ keys, int64_t n, int64_t k) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) poly_masks_nearest_kernel(const int64_t* __restrict__ keys, int64_t n_keys,
                                                                 const int32_t* __restrict__ list_ofs, const int32_t* __restrict__ hw,
                                                                 int out, uint8_t* __restrict__ dst) {
  const int n = blockIdx.y;
  const int h = hw[2 * n], w = hw[2 * n + 1];
  const int l0 = list_ofs[n], l1 = list_ofs[n + 1];
  const int64_t total = (int64_t)out * out;
  uint8_t* d = dst + (int64_t)n * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % out), Y = (int)(i / out);
    const int64_t pos = (int64_t)nearest_idx(X, w, out) * h + nearest_idx(Y, h, out);   // column-major
    int val = 0;
    for (int l = l0; l < l1 && !val; ++l) {          // union of the object's polygons (mask_utils.merge)
      const int64_t base = (int64_t)l << 32;
      const int64_t a = key_lower_bound(keys, n_keys, base), b = key_lower_bound(keys, n_keys, base + pos + 1);
      val = (int)((b - a) & 1);                      // parity of the crossings at or before this position
    }
    d[i] = (uint8_t)val;
  }
}

}  // namespace

int resample_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* coeffs) {
  if (in_size <= 0 || out_size <= 0) return fail(-1, "resample_coeffs: sizes must be positive (in %d, out %d)", in_size, out_size);
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;             // BILINEAR: triangle filter of support 1
  const int ksize = (int)std::ceil(support) * 2 + 1;
  if (coeffs == nullptr || bounds == nullptr) return ksize;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    double k[64];
    const int nk = std::min(ksize, 64);
    for (int x = 0; x < nk; ++x) k[x] = 0.0;
    for (int x = 0; x < xmax && x < nk; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      const double wv = t < 1.0 ? 1.0 - t : 0.0;
      k[x] = wv;
      ww += wv;
    }
    for (int x = 0; x < xmax && x < nk; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < ksize; ++x) {
      const double v = x < nk ? k[x] : 0.0;
      coeffs[(int64_t)xx * ksize + x] = v < 0 ? (int32_t)(-0.5 + v * (1 << PRECISION_BITS)) : (int32_t)(0.5 + v * (1 << PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

int image_resize_normalize(const uint8_t* src, int h, int w, int out, const int32_t* bounds_x, const int32_t* coeffs_x, int ks_x,
                           const int32_t* bounds_y, const int32_t* coeffs_y, int ks_y, uint8_t* tmp, float* dst, float mean,
                           float std, cudaStream_t s) {
  SAM3B_REQUIRE(src && tmp && dst && bounds_x && coeffs_x && bounds_y && coeffs_y, "image_resize_normalize: null tensor");
  SAM3B_REQUIRE(h > 0 && w > 0 && out > 0 && ks_x > 0 && ks_y > 0 && std != 0.f, "image_resize_normalize: bad sizes");
  const int64_t t1 = (int64_t)h * out, t2 = (int64_t)out * out;
  resize_h_kernel<<<(int)std::min<int64_t>((t1 + 255) / 256, (int64_t)num_sms() * 16), 256, 0, s>>>(src, h, w, out, bounds_x, coeffs_x, ks_x, tmp);
  SAM3B_LAUNCHED();
  resize_v_normalize_kernel<<<(int)std::min<int64_t>((t2 + 255) / 256, (int64_t)num_sms() * 16), 256, 0, s>>>(tmp, out, out, bounds_y, coeffs_y, ks_y, mean, std, dst);
  SAM3B_LAUNCHED();
  return 0;
}

int rle_masks_nearest(const uint32_t* cum, const int32_t* offs, const int32_t* hw, int N, int out, uint8_t* dst, cudaStream_t s) {
  if (N <= 0) return 0;
  SAM3B_REQUIRE(cum && offs && hw && dst && out > 0 && N <= 65535, "rle_masks_nearest: bad arguments");
  const int64_t total = (int64_t)out * out;
  const dim3 grid((unsigned)std::min<int64_t>((total + 255) / 256, 1024), N);
  rle_masks_nearest_kernel<<<grid, 256, 0, s>>>(cum, offs, hw, out, dst);
  SAM3B_LAUNCHED();
  return 0;
}

int poly_crossings(const int32_t* edges, const int32_t* pt_start, int n_edges, int64_t total_pts, int64_t* keys, cudaStream_t s) {
  if (total_pts <= 0) return 0;
  SAM3B_REQUIRE(edges && pt_start && keys && n_edges > 0, "poly_crossings: bad arguments");
  const int blocks = (int)std::min<int64_t>((total_pts + 255) / 256, (int64_t)num_sms() * 16);
  poly_crossings_kernel<<<blocks, 256, 0, s>>>(edges, pt_start, n_edges, total_pts, keys);
  SAM3B_LAUNCHED();
  return 0;
}

int poly_masks_nearest(const int64_t* keys_sorted, int64_t n_keys, const int32_t* list_ofs, const int32_t* hw, int N, int out,
                       uint8_t* dst, cudaStream_t s) {
  if (N <= 0) return 0;
  SAM3B_REQUIRE(list_ofs && hw && dst && out > 0 && N <= 65535 && (n_keys == 0 || keys_sorted), "poly_masks_nearest: bad arguments");
  const int64_t total = (int64_t)out * out;
  const dim3 grid((unsigned)std::min<int64_t>((total + 255) / 256, 1024), N);
  poly_masks_nearest_kernel<<<grid, 256, 0, s>>>(keys_sorted, n_keys, list_ofs, hw, out, dst);
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace sam3b
