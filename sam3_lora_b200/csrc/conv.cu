// Neck / pixel-decoder / mask-head helper kernels (see conv.cuh).  All are single-pass, HBM-bound, 128-bit accesses where
// the layout allows; grids are sized from the SM count (grid-stride loops).
#include "conv.cuh"

#include <algorithm>
#include <cmath>

#include "common.h"
#include "gelu.cuh"
#include "ptx.cuh"

namespace sam3b {

namespace {

template <int DT>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack2<DT>(u.x), b = unpack2<DT>(u.y), c = unpack2<DT>(u.z), d = unpack2<DT>(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <int DT>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2<DT>(f[0], f[1]), pack2<DT>(f[2], f[3]), pack2<DT>(f[4], f[5]), pack2<DT>(f[6], f[7]));
}
template <int DT>
__device__ __forceinline__ float load16(const uint16_t* p) {
  const uint32_t v = *p;
  return unpack2<DT>(v).x;
}
template <int DT>
__device__ __forceinline__ void store16(uint16_t* p, float v) { *p = (uint16_t)(pack2<DT>(v, 0.f) & 0xffffu); }

inline int grid_for(int64_t n, int block = 256) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, (int64_t)num_sms() * 16));
}

// ------------------------------------------------------------------------------------------------------------------
// gradient scale
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ g, int64_t n, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(g) & 15) == 0) ? n / 4 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(g[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
__global__ void scale_finalize_kernel(float* scale, float target) {
  const float amax = __uint_as_float(reinterpret_cast<unsigned*>(scale)[2]);
  float s = 1.f;
  if (amax > 0.f && amax < 3.0e38f) {
    int e = (int)floorf(log2f(target / amax));
    e = max(-96, min(96, e));
    s = exp2f((float)e);
  }
  scale[0] = s;
  scale[1] = 1.f / s;
}

// ------------------------------------------------------------------------------------------------------------------
// casts / transposes
// ------------------------------------------------------------------------------------------------------------------
template <int TIN, int TOUT, int DT, bool ACC>
__global__ void __launch_bounds__(256) scale_cast_kernel(const void* __restrict__ in, void* __restrict__ out, int64_t n4,
                                                         const float* __restrict__ scale) {
  const float sc = scale != nullptr ? __ldg(scale) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v;
    if constexpr (TIN == 1) {
      v = reinterpret_cast<const float4*>(in)[i];
    } else {
      const uint2 u = reinterpret_cast<const uint2*>(in)[i];
      const float2 a = unpack2<DT>(u.x), b = unpack2<DT>(u.y);
      v = make_float4(a.x, a.y, b.x, b.y);
    }
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    if constexpr (TOUT == 1) {
      float4* o = reinterpret_cast<float4*>(out) + i;
      if constexpr (ACC) { const float4 t = *o; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
      *o = v;
    } else {
      reinterpret_cast<uint2*>(out)[i] = make_uint2(pack2<DT>(v.x, v.y), pack2<DT>(v.z, v.w));
    }
  }
}

template <int TIN, int TOUT, int DT>
__global__ void __launch_bounds__(256) transpose_kernel(const void* __restrict__ in, void* __restrict__ out, int R, int C,
                                                        const float* __restrict__ scale) {
  __shared__ float tile[32][33];
  const int64_t base = (int64_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + threadIdx.x;
    if (r < R && c < C) {
      const int64_t off = base + (int64_t)r * C + c;
      if constexpr (TIN == 1) tile[k][threadIdx.x] = reinterpret_cast<const float*>(in)[off];
      else tile[k][threadIdx.x] = load16<DT>(reinterpret_cast<const uint16_t*>(in) + off);
    }
  }
  __syncthreads();
  const float sc = scale != nullptr ? __ldg(scale) : 1.f;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      const int64_t off = base + (int64_t)c * R + r;
      const float v = tile[threadIdx.x][k] * sc;
      if constexpr (TOUT == 1) reinterpret_cast<float*>(out)[off] = v;
      else store16<DT>(reinterpret_cast<uint16_t*>(out) + off, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// im2col, pixel shuffle, pooling, upsampling
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col3x3_kernel(const uint4* __restrict__ x, int H, int W, int C8, int64_t total,
                                                        uint16_t* __restrict__ out, int64_t ldo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int tap = (int)(t % 9);
    const int64_t m = t / 9;
    const int px = (int)(m % W);
    const int py = (int)((m / W) % H);
    const int64_t b = m / ((int64_t)W * H);
    const int sy = py + tap / 3 - 1, sx = px + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = x[((b * H + sy) * W + sx) * C8 + c8];
    *reinterpret_cast<uint4*>(out + m * ldo + (int64_t)tap * C8 * 8 + c8 * 8) = v;
  }
}

template <int DT, bool GELU>
__global__ void __launch_bounds__(256) pixel_shuffle2_kernel(const uint4* __restrict__ in, int H, int W, int C8, int64_t total,
                                                             uint4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int X = (int)(t % (2 * W));
    const int Y = (int)((t / (2 * W)) % (2 * H));
    const int64_t b = t / ((int64_t)4 * W * H);
    const int64_t m = (b * H + (Y >> 1)) * W + (X >> 1);
    uint4 v = in[(m * 4 + ((Y & 1) * 2 + (X & 1))) * C8 + c8];
    if constexpr (GELU) {
      float f[8];
      unpack8<DT>(v, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = gelu_erf(f[k]);
      v = pack8<DT>(f);
    }
    out[i] = v;
  }
}

template <int DT, bool DGELU>
__global__ void __launch_bounds__(256) pixel_unshuffle2_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ h, int H,
                                                               int W, int C8, int64_t total, uint4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int blk = (int)(t % 4);
    const int64_t m = t / 4;
    const int j = (int)(m % W);
    const int ii = (int)((m / W) % H);
    const int64_t b = m / ((int64_t)W * H);
    uint4 v = dy[((b * 2 * H + 2 * ii + (blk >> 1)) * 2 * W + 2 * j + (blk & 1)) * C8 + c8];
    if constexpr (DGELU) {
      float f[8], hh[8];
      unpack8<DT>(v, f);
      unpack8<DT>(h[i], hh);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] *= dgelu_erf(hh[k]);
      v = pack8<DT>(f);
    }
    out[i] = v;
  }
}

template <int DT>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const uint4* __restrict__ x, int H, int W, int C8, int64_t total,
                                                           uint4* __restrict__ y) {
  const int Ho = H / 2, Wo = W / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int ox = (int)(t % Wo), oy = (int)((t / Wo) % Ho);
    const int64_t b = t / ((int64_t)Wo * Ho);
    float best[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float f[8];
      unpack8<DT>(x[((b * H + 2 * oy + (p >> 1)) * W + 2 * ox + (p & 1)) * C8 + c8], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) best[k] = (p == 0 || f[k] > best[k]) ? f[k] : best[k];
    }
    y[i] = pack8<DT>(best);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int H, int W,
                                                           int C8, int64_t total, const float* __restrict__ scale,
                                                           float* __restrict__ dx) {
  const int Ho = H / 2, Wo = W / 2;
  const float sc = scale != nullptr ? __ldg(scale) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int ox = (int)(t % Wo), oy = (int)((t / Wo) % Ho);
    const int64_t b = t / ((int64_t)Wo * Ho);
    float best[8], g[8];
    int arg[8];
    unpack8<DT>(dy[i], g);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float f[8];
      unpack8<DT>(x[((b * H + 2 * oy + (p >> 1)) * W + 2 * ox + (p & 1)) * C8 + c8], f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (p == 0 || f[k] > best[k]) { best[k] = f[k]; arg[k] = p; }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float4* d = reinterpret_cast<float4*>(dx + (((b * H + 2 * oy + (p >> 1)) * W + 2 * ox + (p & 1)) * C8 + c8) * 8);
      float4 a = d[0], c = d[1];
      a.x += arg[0] == p ? g[0] * sc : 0.f; a.y += arg[1] == p ? g[1] * sc : 0.f;
      a.z += arg[2] == p ? g[2] * sc : 0.f; a.w += arg[3] == p ? g[3] * sc : 0.f;
      c.x += arg[4] == p ? g[4] * sc : 0.f; c.y += arg[5] == p ? g[5] * sc : 0.f;
      c.z += arg[6] == p ? g[6] * sc : 0.f; c.w += arg[7] == p ? g[7] * sc : 0.f;
      d[0] = a; d[1] = c;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256) upsample_add_kernel(const uint4* __restrict__ prev, int h, int w, const uint4* __restrict__ cur,
                                                           int H, int W, int C8, int64_t total, uint4* __restrict__ out) {
  const int fh = H / h, fw = W / w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int X = (int)(t % W), Y = (int)((t / W) % H);
    const int64_t b = t / ((int64_t)W * H);
    float a[8], p[8];
    unpack8<DT>(cur[i], a);
    unpack8<DT>(prev[((b * h + Y / fh) * w + X / fw) * C8 + c8], p);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += p[k];
    out[i] = pack8<DT>(a);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) upsample_add_bwd_kernel(const uint4* __restrict__ dout, int H, int W, int C8, int h, int w,
                                                               int64_t total, uint4* __restrict__ dprev) {
  const int fh = H / h, fw = W / w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t t = i / C8;
    const int x = (int)(t % w), y = (int)((t / w) % h);
    const int64_t b = t / ((int64_t)w * h);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int dy = 0; dy < fh; ++dy)
      for (int dx = 0; dx < fw; ++dx) {
        float f[8];
        unpack8<DT>(dout[((b * H + y * fh + dy) * W + x * fw + dx) * C8 + c8], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    dprev[i] = pack8<DT>(acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm + ReLU on NHWC fp32
// ------------------------------------------------------------------------------------------------------------------
constexpr int GN_PIX = 512;   // pixels per block of the statistics kernels

// MODE 0: (sum x, sum x^2).  MODE 1: (sum dyh, sum dyh*xhat) with dyh = [y > 0] * dy * gamma.
template <int MODE, int DT>
__global__ void __launch_bounds__(256) gn_sums_kernel(const float4* __restrict__ x, const uint2* __restrict__ dy,
                                                      const float2* __restrict__ stat, const float4* __restrict__ gamma,
                                                      const float4* __restrict__ beta, int HW, int C4, int G, int lanes_per_group,
                                                      double* __restrict__ work) {
  extern __shared__ double sh[];   // [G][2]
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * GN_PIX, p1 = min(HW, p0 + GN_PIX);
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const int q = threadIdx.x % C4, pstep = blockDim.x / C4;
  const int g = (int)(((int64_t)q * G) / C4);   // q*4 / (C/G)
  float s0 = 0.f, s1 = 0.f;
  float4 ga = make_float4(0, 0, 0, 0), be = ga;
  float2 st = make_float2(0, 0);
  if constexpr (MODE == 1) { ga = __ldg(gamma + q); be = __ldg(beta + q); st = __ldg(stat + b * G + g); }
  for (int p = p0 + threadIdx.x / C4; p < p1; p += pstep) {
    const int64_t idx = ((int64_t)b * HW + p) * C4 + q;
    const float4 v = x[idx];
    if constexpr (MODE == 0) {
      s0 += (v.x + v.y) + (v.z + v.w);
      s1 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    } else {
      const uint2 u = dy[idx];
      const float2 d0 = unpack2<DT>(u.x), d1 = unpack2<DT>(u.y);
      const float xh0 = (v.x - st.x) * st.y, xh1 = (v.y - st.x) * st.y, xh2 = (v.z - st.x) * st.y, xh3 = (v.w - st.x) * st.y;
      const float e0 = fmaf(xh0, ga.x, be.x) > 0.f ? d0.x * ga.x : 0.f;
      const float e1 = fmaf(xh1, ga.y, be.y) > 0.f ? d0.y * ga.y : 0.f;
      const float e2 = fmaf(xh2, ga.z, be.z) > 0.f ? d1.x * ga.z : 0.f;
      const float e3 = fmaf(xh3, ga.w, be.w) > 0.f ? d1.y * ga.w : 0.f;
      s0 += (e0 + e1) + (e2 + e3);
      s1 += (e0 * xh0 + e1 * xh1) + (e2 * xh2 + e3 * xh3);
    }
  }
  // lanes of one group are contiguous and aligned when lanes_per_group is a power of two (host-checked): shuffle first
  for (int o = 1; o < lanes_per_group; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & (lanes_per_group - 1)) == 0) {
    atomicAdd(&sh[2 * g], (double)s0);
    atomicAdd(&sh[2 * g + 1], (double)s1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&work[(int64_t)b * 2 * G + i], sh[i]);
}

// MODE 0: stat = (mean, rstd).  MODE 1: out = (sum0 / n, sum1 / n).
template <int MODE>
__global__ void gn_finalize_kernel(const double* __restrict__ work, int BG, double n, float eps, float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BG) return;
  const double a = work[2 * i] / n, c = work[2 * i + 1] / n;
  if constexpr (MODE == 0) {
    const double var = fmax(c - a * a, 0.0);
    out[i] = make_float2((float)a, (float)(1.0 / sqrt(var + (double)eps)));
  } else {
    out[i] = make_float2((float)a, (float)c);
  }
}

template <int DT, bool OUT32>
__global__ void __launch_bounds__(256) gn_relu_fwd_kernel(const float4* __restrict__ x, const float2* __restrict__ stat,
                                                          const float4* __restrict__ gamma, const float4* __restrict__ beta, int HW,
                                                          int C4, int G, int64_t total, void* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    const int b = (int)(i / ((int64_t)HW * C4));
    const int g = (int)(((int64_t)q * G) / C4);
    const float2 st = __ldg(stat + b * G + g);
    const float4 ga = __ldg(gamma + q), be = __ldg(beta + q), v = x[i];
    float4 o;
    o.x = fmaxf(fmaf((v.x - st.x) * st.y, ga.x, be.x), 0.f);
    o.y = fmaxf(fmaf((v.y - st.x) * st.y, ga.y, be.y), 0.f);
    o.z = fmaxf(fmaf((v.z - st.x) * st.y, ga.z, be.z), 0.f);
    o.w = fmaxf(fmaf((v.w - st.x) * st.y, ga.w, be.w), 0.f);
    if constexpr (OUT32) reinterpret_cast<float4*>(y)[i] = o;
    else reinterpret_cast<uint2*>(y)[i] = make_uint2(pack2<DT>(o.x, o.y), pack2<DT>(o.z, o.w));
  }
}

template <int DT>
__global__ void __launch_bounds__(256) gn_relu_bwd_kernel(const uint2* __restrict__ dy, const float4* __restrict__ x,
                                                          const float2* __restrict__ stat, const float2* __restrict__ dmean,
                                                          const float4* __restrict__ gamma, const float4* __restrict__ beta, int HW,
                                                          int C4, int G, int64_t total, uint2* __restrict__ dx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    const int b = (int)(i / ((int64_t)HW * C4));
    const int g = (int)(((int64_t)q * G) / C4);
    const float2 st = __ldg(stat + b * G + g), dm = __ldg(dmean + b * G + g);
    const float4 ga = __ldg(gamma + q), be = __ldg(beta + q), v = x[i];
    const uint2 u = dy[i];
    const float2 d0 = unpack2<DT>(u.x), d1 = unpack2<DT>(u.y);
    const float xh0 = (v.x - st.x) * st.y, xh1 = (v.y - st.x) * st.y, xh2 = (v.z - st.x) * st.y, xh3 = (v.w - st.x) * st.y;
    const float e0 = fmaf(xh0, ga.x, be.x) > 0.f ? d0.x * ga.x : 0.f;
    const float e1 = fmaf(xh1, ga.y, be.y) > 0.f ? d0.y * ga.y : 0.f;
    const float e2 = fmaf(xh2, ga.z, be.z) > 0.f ? d1.x * ga.z : 0.f;
    const float e3 = fmaf(xh3, ga.w, be.w) > 0.f ? d1.y * ga.w : 0.f;
    const float r0 = st.y * (e0 - dm.x - xh0 * dm.y), r1 = st.y * (e1 - dm.x - xh1 * dm.y);
    const float r2 = st.y * (e2 - dm.x - xh2 * dm.y), r3 = st.y * (e3 - dm.x - xh3 * dm.y);
    dx[i] = make_uint2(pack2<DT>(r0, r1), pack2<DT>(r2, r3));
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

#define SAM3B_DT_SWITCH(dtype, ...)                         \
  do {                                                      \
    if ((dtype) == 0) { constexpr int DT = 0; __VA_ARGS__; } \
    else { constexpr int DT = 1; __VA_ARGS__; }             \
  } while (0)

int grad_scale(const float* g, int64_t n, float target, float* scale, cudaStream_t s) {
  SAM3B_REQUIRE(g && scale && n > 0 && target > 0.f, "grad_scale: bad arguments");
  SAM3B_CHECK_CUDA(cudaMemsetAsync(scale + 2, 0, sizeof(float), s));
  amax_kernel<<<grid_for(n / 4 + 1), 256, 0, s>>>(g, n, reinterpret_cast<unsigned*>(scale) + 2);
  SAM3B_LAUNCHED();
  scale_finalize_kernel<<<1, 1, 0, s>>>(scale, target);
  SAM3B_LAUNCHED();
  return 0;
}

int scale_cast(const void* in, int tin, void* out, int tout, int64_t n, int dtype, const float* scale, int accumulate,
               cudaStream_t s) {
  if (n <= 0) return 0;
  SAM3B_REQUIRE(in && out && n % 4 == 0, "scale_cast: null tensor or n %% 4 != 0");
  SAM3B_REQUIRE(aligned16(in) && aligned16(out), "scale_cast: tensors must be 16-byte aligned");
  SAM3B_REQUIRE(!(accumulate && tout != 1), "scale_cast: accumulate needs an fp32 destination");
  const int64_t n4 = n / 4;
  const int blocks = grid_for(n4);
  SAM3B_DT_SWITCH(dtype, {
    if (tin == 1 && tout == 0) scale_cast_kernel<1, 0, DT, false><<<blocks, 256, 0, s>>>(in, out, n4, scale);
    else if (tin == 0 && tout == 1 && accumulate) scale_cast_kernel<0, 1, DT, true><<<blocks, 256, 0, s>>>(in, out, n4, scale);
    else if (tin == 0 && tout == 1) scale_cast_kernel<0, 1, DT, false><<<blocks, 256, 0, s>>>(in, out, n4, scale);
    else if (tin == 1 && tout == 1 && accumulate) scale_cast_kernel<1, 1, DT, true><<<blocks, 256, 0, s>>>(in, out, n4, scale);
    else if (tin == 1 && tout == 1) scale_cast_kernel<1, 1, DT, false><<<blocks, 256, 0, s>>>(in, out, n4, scale);
    else return fail(-1, "scale_cast: unsupported type pair %d -> %d", tin, tout);
  });
  SAM3B_LAUNCHED();
  return 0;
}

int transpose_cast(const void* in, int tin, void* out, int tout, int batch, int R, int C, int dtype, const float* scale,
                   cudaStream_t s) {
  if (batch <= 0 || R <= 0 || C <= 0) return 0;
  SAM3B_REQUIRE(in && out, "transpose_cast: null tensor");
  SAM3B_REQUIRE(batch <= 65535 && (R + 31) / 32 <= 65535, "transpose_cast: grid too large (batch %d, R %d)", batch, R);
  const dim3 grid((C + 31) / 32, (R + 31) / 32, batch), block(32, 8);
  SAM3B_DT_SWITCH(dtype, {
    if (tin == 1 && tout == 0) transpose_kernel<1, 0, DT><<<grid, block, 0, s>>>(in, out, R, C, scale);
    else if (tin == 0 && tout == 1) transpose_kernel<0, 1, DT><<<grid, block, 0, s>>>(in, out, R, C, scale);
    else if (tin == 0 && tout == 0) transpose_kernel<0, 0, DT><<<grid, block, 0, s>>>(in, out, R, C, scale);
    else transpose_kernel<1, 1, DT><<<grid, block, 0, s>>>(in, out, R, C, scale);
  });
  SAM3B_LAUNCHED();
  return 0;
}

int im2col3x3(const void* x16, int B, int H, int W, int C, void* out16, int64_t ldo, cudaStream_t s) {
  SAM3B_REQUIRE(x16 && out16 && B > 0 && H > 0 && W > 0, "im2col3x3: bad arguments");
  SAM3B_REQUIRE(C % 8 == 0 && ldo % 8 == 0 && ldo >= 9 * (int64_t)C, "im2col3x3: C %% 8, ldo %% 8, ldo >= 9C (C=%d ldo=%lld)", C,
                (long long)ldo);
  SAM3B_REQUIRE(aligned16(x16) && aligned16(out16), "im2col3x3: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * H * W * 9 * (C / 8);
  im2col3x3_kernel<<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(x16), H, W, C / 8, total,
                                                  reinterpret_cast<uint16_t*>(out16), ldo);
  SAM3B_LAUNCHED();
  return 0;
}

int pixel_shuffle2(const void* in16, int B, int H, int W, int C, int gelu, void* out16, int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(in16 && out16 && B > 0 && H > 0 && W > 0 && C % 8 == 0, "pixel_shuffle2: bad arguments (C=%d)", C);
  SAM3B_REQUIRE(aligned16(in16) && aligned16(out16), "pixel_shuffle2: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
  const int blocks = grid_for(total);
  SAM3B_DT_SWITCH(dtype, {
    if (gelu) pixel_shuffle2_kernel<DT, true><<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(in16), H, W, C / 8, total, reinterpret_cast<uint4*>(out16));
    else pixel_shuffle2_kernel<DT, false><<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(in16), H, W, C / 8, total, reinterpret_cast<uint4*>(out16));
  });
  SAM3B_LAUNCHED();
  return 0;
}

int pixel_unshuffle2(const void* dy16, const void* h16, int B, int H, int W, int C, void* out16, int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(dy16 && out16 && B > 0 && H > 0 && W > 0 && C % 8 == 0, "pixel_unshuffle2: bad arguments (C=%d)", C);
  SAM3B_REQUIRE(aligned16(dy16) && aligned16(out16) && aligned16(h16), "pixel_unshuffle2: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
  const int blocks = grid_for(total);
  SAM3B_DT_SWITCH(dtype, {
    if (h16) pixel_unshuffle2_kernel<DT, true><<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(dy16), reinterpret_cast<const uint4*>(h16), H, W, C / 8, total, reinterpret_cast<uint4*>(out16));
    else pixel_unshuffle2_kernel<DT, false><<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(dy16), nullptr, H, W, C / 8, total, reinterpret_cast<uint4*>(out16));
  });
  SAM3B_LAUNCHED();
  return 0;
}

int maxpool2_fwd(const void* x16, int B, int H, int W, int C, void* y16, int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(x16 && y16 && B > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool2_fwd: even H, W and C %% 8 expected");
  SAM3B_REQUIRE(aligned16(x16) && aligned16(y16), "maxpool2_fwd: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
  SAM3B_DT_SWITCH(dtype, maxpool2_fwd_kernel<DT><<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(x16), H, W, C / 8, total, reinterpret_cast<uint4*>(y16)));
  SAM3B_LAUNCHED();
  return 0;
}

int maxpool2_bwd(const void* x16, const void* dy16, int B, int H, int W, int C, const float* scale, float* dx32, int dtype,
                 cudaStream_t s) {
  SAM3B_REQUIRE(x16 && dy16 && dx32 && B > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool2_bwd: even H, W and C %% 8 expected");
  SAM3B_REQUIRE(aligned16(x16) && aligned16(dy16) && aligned16(dx32), "maxpool2_bwd: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
  SAM3B_DT_SWITCH(dtype, maxpool2_bwd_kernel<DT><<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(x16), reinterpret_cast<const uint4*>(dy16), H, W, C / 8, total, scale, dx32));
  SAM3B_LAUNCHED();
  return 0;
}

int upsample_add(const void* prev16, int h, int w, const void* cur16, int B, int H, int W, int C, void* out16, int dtype,
                 cudaStream_t s) {
  SAM3B_REQUIRE(prev16 && cur16 && out16 && B > 0 && C % 8 == 0, "upsample_add: bad arguments");
  SAM3B_REQUIRE(h > 0 && w > 0 && H % h == 0 && W % w == 0, "upsample_add: %dx%d is not an integer multiple of %dx%d", H, W, h, w);
  SAM3B_REQUIRE(aligned16(prev16) && aligned16(cur16) && aligned16(out16), "upsample_add: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * H * W * (C / 8);
  SAM3B_DT_SWITCH(dtype, upsample_add_kernel<DT><<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(prev16), h, w, reinterpret_cast<const uint4*>(cur16), H, W, C / 8, total, reinterpret_cast<uint4*>(out16)));
  SAM3B_LAUNCHED();
  return 0;
}

int upsample_add_bwd(const void* dout16, int B, int H, int W, int C, int h, int w, void* dprev16, int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(dout16 && dprev16 && B > 0 && C % 8 == 0, "upsample_add_bwd: bad arguments");
  SAM3B_REQUIRE(h > 0 && w > 0 && H % h == 0 && W % w == 0, "upsample_add_bwd: %dx%d is not an integer multiple of %dx%d", H, W, h, w);
  SAM3B_REQUIRE(aligned16(dout16) && aligned16(dprev16), "upsample_add_bwd: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)B * h * w * (C / 8);
  SAM3B_DT_SWITCH(dtype, upsample_add_bwd_kernel<DT><<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(dout16), H, W, C / 8, h, w, total, reinterpret_cast<uint4*>(dprev16)));
  SAM3B_LAUNCHED();
  return 0;
}

// threads per block and lanes per group for the statistics kernels; fails for channel counts the mapping cannot take
static int gn_geometry(int C, int G, int* threads, int* lanes_per_group) {
  SAM3B_REQUIRE(C > 0 && G > 0 && C % G == 0 && (C / G) % 4 == 0, "groupnorm: C=%d must split into %d groups of a multiple of 4 channels", C, G);
  const int C4 = C / 4;
  SAM3B_REQUIRE(C4 <= 256 && (C4 % 32 == 0 || 32 % C4 == 0), "groupnorm: C/4=%d must divide or be a multiple of 32 (<= 256)", C4);
  *threads = C4 * (256 / C4);
  const int l = (C / G) / 4;
  *lanes_per_group = ((l & (l - 1)) == 0 && l <= 32) ? l : 1;
  return 0;
}

int groupnorm_stats(const float* x, int B, int HW, int C, int G, float eps, double* work, float* stat, cudaStream_t s) {
  SAM3B_REQUIRE(x && work && stat && B > 0 && HW > 0, "groupnorm_stats: bad arguments");
  int threads, lpg, rc;
  if ((rc = gn_geometry(C, G, &threads, &lpg))) return rc;
  SAM3B_CHECK_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * 2 * B * G, s));
  const dim3 grid((HW + GN_PIX - 1) / GN_PIX, B);
  gn_sums_kernel<0, 0><<<grid, threads, 2 * G * sizeof(double), s>>>(reinterpret_cast<const float4*>(x), nullptr, nullptr, nullptr, nullptr, HW, C / 4, G, lpg, work);
  SAM3B_LAUNCHED();
  gn_finalize_kernel<0><<<(B * G + 127) / 128, 128, 0, s>>>(work, B * G, (double)HW * (C / G), eps, reinterpret_cast<float2*>(stat));
  SAM3B_LAUNCHED();
  return 0;
}

int groupnorm_relu_fwd(const float* x, const float* stat, const float* gamma, const float* beta, int B, int HW, int C, int G,
                       void* y, int out_f32, int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(x && stat && gamma && beta && y && B > 0 && HW > 0, "groupnorm_relu_fwd: bad arguments");
  int threads, lpg, rc;
  if ((rc = gn_geometry(C, G, &threads, &lpg))) return rc;
  const int64_t total = (int64_t)B * HW * (C / 4);
  const int blocks = grid_for(total);
  SAM3B_DT_SWITCH(dtype, {
    if (out_f32) gn_relu_fwd_kernel<DT, true><<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float2*>(stat), reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta), HW, C / 4, G, total, y);
    else gn_relu_fwd_kernel<DT, false><<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float2*>(stat), reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta), HW, C / 4, G, total, y);
  });
  SAM3B_LAUNCHED();
  return 0;
}

int groupnorm_relu_bwd(const void* dy16, const float* x, const float* stat, const float* gamma, const float* beta, int B, int HW,
                       int C, int G, double* work, void* dx16, int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(dy16 && x && stat && gamma && beta && work && dx16 && B > 0 && HW > 0, "groupnorm_relu_bwd: bad arguments");
  int threads, lpg, rc;
  if ((rc = gn_geometry(C, G, &threads, &lpg))) return rc;
  SAM3B_CHECK_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * 2 * B * G, s));
  float2* dmean = reinterpret_cast<float2*>(work + 2 * (int64_t)B * G);   // third B*G doubles of `work`
  const dim3 grid((HW + GN_PIX - 1) / GN_PIX, B);
  const int64_t total = (int64_t)B * HW * (C / 4);
  SAM3B_DT_SWITCH(dtype, {
    gn_sums_kernel<1, DT><<<grid, threads, 2 * G * sizeof(double), s>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const uint2*>(dy16), reinterpret_cast<const float2*>(stat), reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta), HW, C / 4, G, lpg, work);
    SAM3B_LAUNCHED();
    gn_finalize_kernel<1><<<(B * G + 127) / 128, 128, 0, s>>>(work, B * G, (double)HW * (C / G), 0.f, dmean);
    SAM3B_LAUNCHED();
    gn_relu_bwd_kernel<DT><<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint2*>(dy16), reinterpret_cast<const float4*>(x), reinterpret_cast<const float2*>(stat), dmean, reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta), HW, C / 4, G, total, reinterpret_cast<uint2*>(dx16));
  });
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace sam3b
