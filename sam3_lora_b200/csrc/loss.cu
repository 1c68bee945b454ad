// Fused sigmoid focal loss forward / backward (next-row f1 of the hot-path table; replaces the reference's
// Triton kernels sam3/train/loss/sigmoid_focal_loss.py:35-208, same arithmetic as the `triton=False` branch of
// sam3/train/loss/loss_fns.py:159-167):
//   p = sigmoid(x);  bce = max(x,0) - x*y + log1p(exp(-|x|));  p_t = p*y + (1-p)*(1-y)
//   L = [alpha*y + (1-alpha)*(1-y)] * (1 - p_t)^gamma * bce          (alpha < 0: no alpha factor)
// HBM-bound elementwise kernels: 128-bit loads, one pass; the reduced variant adds a warp-shuffle + one atomic
// per block.  The backward recomputes everything from (x, y) so nothing but the inputs is saved.
#include "loss.cuh"

#include <algorithm>

#include "common.h"

namespace sam3b {

namespace {

__device__ __forceinline__ void focal_terms(float x, float y, float alpha, float gamma, float& loss, float& dldx) {
  const float e = __expf(-fabsf(x));
  const float inv = 1.f / (1.f + e);
  const float p = x >= 0.f ? inv : e * inv;                 // sigmoid(x), stable on both sides
  const float bce = fmaxf(x, 0.f) - x * y + log1pf(e);
  const float s = 2.f * y - 1.f;
  const float u = y - p * s;                                 // 1 - p_t
  const float at = alpha >= 0.f ? alpha * y + (1.f - alpha) * (1.f - y) : 1.f;
  float mod, dmod;                                           // u^gamma and d(u^gamma)/du
  if (gamma == 2.f) { mod = u * u; dmod = 2.f * u; }
  else if (gamma == 0.f) { mod = 1.f; dmod = 0.f; }
  else { const float um = fmaxf(u, 1e-30f); mod = __powf(um, gamma); dmod = gamma * __powf(um, gamma - 1.f); }
  loss = at * mod * bce;
  const float dudx = -s * p * (1.f - p);
  dldx = at * (dmod * dudx * bce + mod * (p - y));
}

__global__ void __launch_bounds__(256) focal_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n,
                                                        float alpha, float gamma, float* __restrict__ loss,
                                                        float* __restrict__ sum) {
  float acc = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i], yv = reinterpret_cast<const float4*>(y)[i];
    float4 l; float d;
    focal_terms(xv.x, yv.x, alpha, gamma, l.x, d); focal_terms(xv.y, yv.y, alpha, gamma, l.y, d);
    focal_terms(xv.z, yv.z, alpha, gamma, l.z, d); focal_terms(xv.w, yv.w, alpha, gamma, l.w, d);
    if (loss != nullptr) reinterpret_cast<float4*>(loss)[i] = l;
    acc += (l.x + l.y) + (l.z + l.w);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float l, d;
    focal_terms(x[i], y[i], alpha, gamma, l, d);
    if (loss != nullptr) loss[i] = l;
    acc += l;
  }
  if (sum != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += part[w];
      atomicAdd(sum, t);
    }
  }
}

// dx[i] = dldx(x,y) * (g != null ? g[i] : 1) * gscale
__global__ void __launch_bounds__(256) focal_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n,
                                                        float alpha, float gamma, const float* __restrict__ g, float gscale,
                                                        float* __restrict__ dx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float l, d;
    focal_terms(x[i], y[i], alpha, gamma, l, d);
    dx[i] = d * gscale * (g != nullptr ? g[i] : 1.f);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Fused mask losses (row f1): bilinear up-sampling of the matched mask logits to the target resolution
// (F.interpolate(mode="bilinear", align_corners=False), loss_fns.py:689-696) + sigmoid focal loss + dice loss
// (loss_fns.py:105-123, 126-176, 698-707) in ONE pass over the targets.  The up-sampled fp32 logits (4 MB per mask at
// 1008^2) are never written; the targets may be uint8 (1 B/pixel) or fp32.  HBM-bound on the target read.
//   x_p   = bilerp(src_n, p)                     p over H x W
//   focal = sum_p L(x_p, t_p)                     -> loss_mask = sum_n focal_n / (H W) / num_boxes
//   I = sum_p sig(x_p) t_p, S = sum_p sig(x_p), T = sum_p t_p -> loss_dice = sum_n [1 - (2 I + 1)/(S + T + 1)] / num_boxes
// Deterministic: per-strip partial sums, reduced in a fixed order by the finalize kernel; the backward is a gather over
// the (at most ceil(2 H/h) x ceil(2 W/w)) target pixels whose bilinear stencil touches a source pixel (no atomics).
// ------------------------------------------------------------------------------------------------------------------
constexpr int ML_ROWS = 8;   // target rows per block of the forward pass

// One evaluation per target pixel, sharing exp / reciprocal between the focal and dice terms (these kernels are bound by
// instruction issue, not HBM: 1 byte of target per ~40 instructions).  LOSS: focal loss value; GRAD: d focal / dx.
template <bool LOSS, bool GRAD>
__device__ __forceinline__ void mask_pixel(float x, float t, float alpha, float gamma, float& loss, float& dldx, float& p) {
  const float e = __expf(-fabsf(x));
  const float inv = __fdividef(1.f, 1.f + e);
  p = x >= 0.f ? inv : e * inv;
  const float bce = fmaxf(x, 0.f) - x * t + __logf(1.f + e);
  const float s = 2.f * t - 1.f;
  const float u = t - p * s;                                 // 1 - p_t
  const float at = alpha >= 0.f ? alpha * t + (1.f - alpha) * (1.f - t) : 1.f;
  float mod, dmod;
  if (gamma == 2.f) { mod = u * u; dmod = 2.f * u; }
  else if (gamma == 0.f) { mod = 1.f; dmod = 0.f; }
  else { const float um = fmaxf(u, 1e-30f); mod = __powf(um, gamma); dmod = gamma * __powf(um, gamma - 1.f); }
  if constexpr (LOSS) loss = at * mod * bce;
  if constexpr (GRAD) dldx = at * (dmod * (-s * p * (1.f - p)) * bce + mod * (p - t));
}

struct Lerp { int i0, i1; float l; };
// ATen area_pixel_compute_source_index (align_corners=False, bilinear): src = max(scale*(dst+0.5)-0.5, 0)
__device__ __forceinline__ Lerp lerp_index(int dst, float scale, int in_size) {
  const float s = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  Lerp r;
  r.i0 = min((int)s, in_size - 1);
  r.i1 = min(r.i0 + 1, in_size - 1);
  r.l = s - (float)r.i0;
  return r;
}
template <typename TT> __device__ __forceinline__ float tgt_at(const TT* t, int64_t i) { return (float)t[i]; }

template <typename TT>
__global__ void __launch_bounds__(256) mask_loss_fwd_kernel(const float* __restrict__ src, int h, int w, const TT* __restrict__ tgt,
                                                            int H, int W, float alpha, float gamma, float* __restrict__ partial) {
  const int n = blockIdx.y, strips = gridDim.x;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float* sn = src + (int64_t)n * h * w;
  const TT* tn = tgt + (int64_t)n * H * W;
  float a_f = 0.f, a_i = 0.f, a_s = 0.f, a_t = 0.f;
  const int y_lo = blockIdx.x * ML_ROWS, y_hi = min(H, y_lo + ML_ROWS);
  for (int Y = y_lo; Y < y_hi; ++Y) {
    const Lerp ly = lerp_index(Y, sy, h);
    const float* r0 = sn + (int64_t)ly.i0 * w;
    const float* r1 = sn + (int64_t)ly.i1 * w;
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
      const Lerp lx = lerp_index(X, sx, w);
      const float top = r0[lx.i0] + lx.l * (r0[lx.i1] - r0[lx.i0]);
      const float bot = r1[lx.i0] + lx.l * (r1[lx.i1] - r1[lx.i0]);
      const float x = top + ly.l * (bot - top);
      const float t = tgt_at(tn, (int64_t)Y * W + X);
      float l, d, p;
      mask_pixel<true, false>(x, t, alpha, gamma, l, d, p);
      a_f += l; a_i += p * t; a_s += p; a_t += t;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a_f += __shfl_xor_sync(0xffffffffu, a_f, o); a_i += __shfl_xor_sync(0xffffffffu, a_i, o);
    a_s += __shfl_xor_sync(0xffffffffu, a_s, o); a_t += __shfl_xor_sync(0xffffffffu, a_t, o);
  }
  __shared__ float4 part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = make_float4(a_f, a_i, a_s, a_t);
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 t = part[0];
    for (int k = 1; k < 8; ++k) { t.x += part[k].x; t.y += part[k].y; t.z += part[k].z; t.w += part[k].w; }
    reinterpret_cast<float4*>(partial)[(int64_t)n * strips + blockIdx.x] = t;
  }
}

// one warp per mask: strips reduced in a fixed order in double; thread 0 of block 0 then folds the masks (fixed order)
__global__ void __launch_bounds__(256) mask_loss_finalize_kernel(const float* __restrict__ partial, int N, int strips, double hw,
                                                                 float num_boxes, float* __restrict__ sums, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n = wid; n < N; n += nw) {
    double f = 0, i = 0, s = 0, t = 0;
    for (int k = lane; k < strips; k += 32) {
      const float4 v = reinterpret_cast<const float4*>(partial)[(int64_t)n * strips + k];
      f += v.x; i += v.y; s += v.z; t += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      f += __shfl_xor_sync(0xffffffffu, f, o); i += __shfl_xor_sync(0xffffffffu, i, o);
      s += __shfl_xor_sync(0xffffffffu, s, o); t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (lane == 0) { sums[4 * n] = (float)f; sums[4 * n + 1] = (float)i; sums[4 * n + 2] = (float)s; sums[4 * n + 3] = (float)t; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double lm = 0, ld = 0;
    for (int n = 0; n < N; ++n) {
      lm += (double)sums[4 * n] / hw;
      ld += 1.0 - (2.0 * sums[4 * n + 1] + 1.0) / ((double)sums[4 * n + 2] + (double)sums[4 * n + 3] + 1.0);
    }
    out[0] = (float)(lm / num_boxes);
    out[1] = (float)(ld / num_boxes);
  }
}

// dsrc[n][y][x] = sum over target pixels p whose stencil contains (y, x) of w_p(y, x) * dL/dx_p, with
// dL/dx_p = g[0] / (H W num_boxes) * dfocal + g[1] / num_boxes * ddice_n,  ddice_n/dx_p = -sig'(x_p) [2 t_p (D+1) - (2I+1)] / (D+1)^2
template <typename TT>
__global__ void __launch_bounds__(256) mask_loss_bwd_kernel(const float* __restrict__ src, int h, int w, const TT* __restrict__ tgt,
                                                            int H, int W, float alpha, float gamma, const float* __restrict__ sums,
                                                            const float* __restrict__ g, float num_boxes, int64_t total,
                                                            float* __restrict__ dsrc) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float gm = __ldg(g) / ((float)H * (float)W * num_boxes), gd = __ldg(g + 1) / num_boxes;
  const int ry = (int)ceilf((float)H / (float)h) + 1, rx = (int)ceilf((float)W / (float)w) + 1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w), y = (int)((idx / w) % h);
    const int n = (int)(idx / ((int64_t)w * h));
    const float* sn = src + (int64_t)n * h * w;
    const TT* tn = tgt + (int64_t)n * H * W;
    const float I2 = 2.f * sums[4 * n + 1] + 1.f, D1 = sums[4 * n + 2] + sums[4 * n + 3] + 1.f;
    const float invD2 = 1.f / (D1 * D1);
    // candidate target rows / columns: centres within one source pixel of (y, x)
    const int Yc = (int)(((float)y + 0.5f) / sy), Xc = (int)(((float)x + 0.5f) / sx);
    float acc = 0.f;
    for (int Y = max(0, Yc - ry); Y <= min(H - 1, Yc + ry); ++Y) {
      const Lerp ly = lerp_index(Y, sy, h);
      const float wy = (ly.i0 == y ? 1.f - ly.l : 0.f) + (ly.i1 == y ? ly.l : 0.f);
      if (wy == 0.f) continue;
      const float* r0 = sn + (int64_t)ly.i0 * w;
      const float* r1 = sn + (int64_t)ly.i1 * w;
      for (int X = max(0, Xc - rx); X <= min(W - 1, Xc + rx); ++X) {
        const Lerp lx = lerp_index(X, sx, w);
        const float wx = (lx.i0 == x ? 1.f - lx.l : 0.f) + (lx.i1 == x ? lx.l : 0.f);
        if (wx == 0.f) continue;
        const float top = r0[lx.i0] + lx.l * (r0[lx.i1] - r0[lx.i0]);
        const float bot = r1[lx.i0] + lx.l * (r1[lx.i1] - r1[lx.i0]);
        const float xv = top + ly.l * (bot - top);
        const float t = tgt_at(tn, (int64_t)Y * W + X);
        float l, d, p;
        mask_pixel<false, true>(xv, t, alpha, gamma, l, d, p);
        const float dd = -p * (1.f - p) * (2.f * t * D1 - I2) * invD2;
        acc += wy * wx * (gm * d + gd * dd);
      }
    }
    dsrc[idx] = acc;
  }
}

}  // namespace

int focal_loss_fwd(const float* x, const float* y, int64_t n, float alpha, float gamma, float* loss, float* sum, cudaStream_t s) {
  if (n <= 0) return 0;
  SAM3B_REQUIRE(x && y && (loss || sum), "focal_loss_fwd: null tensor");
  SAM3B_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(loss)) & 15) == 0,
                "focal_loss_fwd: tensors must be 16-byte aligned");
  if (sum) SAM3B_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(float), s));
  const int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)num_sms() * 8);
  focal_fwd_kernel<<<blocks, 256, 0, s>>>(x, y, n, alpha, gamma, loss, sum);
  SAM3B_LAUNCHED();
  return 0;
}

int focal_loss_bwd(const float* x, const float* y, int64_t n, float alpha, float gamma, const float* g, float gscale, float* dx,
                   cudaStream_t s) {
  if (n <= 0) return 0;
  SAM3B_REQUIRE(x && y && dx, "focal_loss_bwd: null tensor");
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8);
  focal_bwd_kernel<<<blocks, 256, 0, s>>>(x, y, n, alpha, gamma, g, gscale, dx);
  SAM3B_LAUNCHED();
  return 0;
}

int mask_loss_fwd(const float* src, int N, int h, int w, const void* tgt, int tgt_u8, int H, int W, float alpha, float gamma,
                  float num_boxes, float* partial, float* sums, float* out, cudaStream_t s) {
  SAM3B_REQUIRE(sums && out && N >= 0 && h > 0 && w > 0 && H > 0 && W > 0 && num_boxes > 0.f, "mask_loss_fwd: bad arguments");
  const int strips = (H + ML_ROWS - 1) / ML_ROWS;
  if (N > 0) {
    SAM3B_REQUIRE(src && tgt && partial, "mask_loss_fwd: null tensor");
    SAM3B_REQUIRE(N <= 65535, "mask_loss_fwd: at most 65535 masks per call (got %d)", N);
    SAM3B_REQUIRE((reinterpret_cast<uintptr_t>(partial) & 15) == 0, "mask_loss_fwd: workspace must be 16-byte aligned");
    const dim3 grid(strips, N);
    if (tgt_u8) mask_loss_fwd_kernel<uint8_t><<<grid, 256, 0, s>>>(src, h, w, static_cast<const uint8_t*>(tgt), H, W, alpha, gamma, partial);
    else mask_loss_fwd_kernel<float><<<grid, 256, 0, s>>>(src, h, w, static_cast<const float*>(tgt), H, W, alpha, gamma, partial);
    SAM3B_LAUNCHED();
  }
  mask_loss_finalize_kernel<<<1, 256, 0, s>>>(partial, N, strips, (double)H * W, num_boxes, sums, out);
  SAM3B_LAUNCHED();
  return 0;
}

int mask_loss_bwd(const float* src, int N, int h, int w, const void* tgt, int tgt_u8, int H, int W, float alpha, float gamma,
                  float num_boxes, const float* sums, const float* g, float* dsrc, cudaStream_t s) {
  if (N <= 0) return 0;
  SAM3B_REQUIRE(src && tgt && sums && g && dsrc && h > 0 && w > 0 && H > 0 && W > 0 && num_boxes > 0.f, "mask_loss_bwd: bad arguments");
  const int64_t total = (int64_t)N * h * w;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
  if (tgt_u8) mask_loss_bwd_kernel<uint8_t><<<blocks, 256, 0, s>>>(src, h, w, static_cast<const uint8_t*>(tgt), H, W, alpha, gamma, sums, g, num_boxes, total, dsrc);
  else mask_loss_bwd_kernel<float><<<blocks, 256, 0, s>>>(src, h, w, static_cast<const float*>(tgt), H, W, alpha, gamma, sums, g, num_boxes, total, dsrc);
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace sam3b
