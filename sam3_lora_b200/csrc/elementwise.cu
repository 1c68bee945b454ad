// HBM-bound kernels of the ViT trunk: LayerNorm fwd/bwd (warp-shuffle reductions, one warp per
// token row, 128-bit accesses), casts, attention delta, patch gather, layout transposes, LoRA
// operand packing, fused AdamW.  All are bandwidth kernels: grids are sized in rows, each thread
// keeps several 16-byte loads in flight; no shared-memory staging is needed except the transposes.
#include "elementwise.cuh"

#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"

#include <algorithm>
#include <type_traits>

namespace sam3b {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DT>
__device__ __forceinline__ uint16_t to16(float v) {
  if constexpr (DT == 0) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  else { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
}

template <int DT>
__device__ __forceinline__ uint2 pack4(float a, float b, float c, float d) {
  uint2 u;
  u.x = pack2<DT>(a, b);
  u.y = pack2<DT>(c, d);
  return u;
}

// ------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, VPT float4 per lane (D = 128*VPT)
// ------------------------------------------------------------------------------------------
template <int VPT, int DT>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int rows,
                                                     uint16_t* __restrict__ y, int64_t ldy, float* __restrict__ mean,
                                                     float* __restrict__ rstd) {
  constexpr int D = 128 * VPT;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * D);
  float4 v[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    v[i] = xr[lane + i * 32];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rs = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + (int64_t)row * ldy);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const float4 g = __ldg(g4 + lane + i * 32), b = __ldg(b4 + lane + i * 32);
    yr[lane + i * 32] = pack4<DT>((v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y,
                                  (v[i].z - mu) * rs * g.z + b.z, (v[i].w - mu) * rs * g.w + b.w);
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward (+ residual-gradient add): dx = dres + rstd*(g - mean(g) - xhat*mean(g*xhat))
// ------------------------------------------------------------------------------------------
template <int VPT, int DT>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const uint16_t* __restrict__ dy, int64_t lddy,
                                                     const float* __restrict__ x, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                     const float* dres, int rows, float* dx, uint16_t* dx16,
                                                     int64_t lddx16, const float* __restrict__ row_scale,
                                                     int rows_per_scale) {
  constexpr int D = 128 * VPT;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (row >= rows) return;
  const float mu = mean[row], rs = rstd[row];
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * D);
  const uint2* dyr = reinterpret_cast<const uint2*>(dy + (int64_t)row * lddy);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  // all three streams (x, dy, residual gradient) are requested before the first reduction: one memory-latency phase per row
  const float4* rr = reinterpret_cast<const float4*>(dres + (int64_t)row * D);
  float4 xh[VPT], g[VPT], rres[VPT];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) rres[i] = rr[lane + i * 32];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const float4 xv = xr[lane + i * 32];
    const uint2 d = dyr[lane + i * 32];
    const float2 d0 = unpack2<DT>(d.x), d1 = unpack2<DT>(d.y);
    const float4 gm = __ldg(g4 + lane + i * 32);
    xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
    g[i] = make_float4(d0.x * gm.x, d0.y * gm.y, d1.x * gm.z, d1.y * gm.w);
    s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
  }
  const float m1 = warp_sum(s1) * (1.f / D), m2 = warp_sum(s2) * (1.f / D);
  float4* dxr = reinterpret_cast<float4*>(dx + (int64_t)row * D);
  const float sc = row_scale != nullptr ? __ldg(row_scale + row / rows_per_scale) : 1.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const float4 r = rres[i];
    float4 o;
    o.x = r.x + rs * (g[i].x - m1 - xh[i].x * m2);
    o.y = r.y + rs * (g[i].y - m1 - xh[i].y * m2);
    o.z = r.z + rs * (g[i].z - m1 - xh[i].z * m2);
    o.w = r.w + rs * (g[i].w - m1 - xh[i].w * m2);
    dxr[lane + i * 32] = o;
    if (dx16 != nullptr)
      reinterpret_cast<uint2*>(dx16 + (int64_t)row * lddx16)[lane + i * 32] = pack4<DT>(o.x * sc, o.y * sc, o.z * sc, o.w * sc);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ x, int rows, int D,
                                                        uint16_t* __restrict__ y, int64_t ldy, const float* __restrict__ scale) {
  const int64_t n4 = (int64_t)rows * (D / 4);
  const float sc = scale != nullptr ? __ldg(scale) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / (D / 4)), c4 = (int)(i % (D / 4));
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    reinterpret_cast<uint2*>(y + (int64_t)row * ldy)[c4] = pack4<DT>(v.x, v.y, v.z, v.w);
  }
}

// delta: 8 lanes per (row, head): each lane covers 8 of the 64 columns with one 16-byte load
template <int DT>
__global__ void __launch_bounds__(256) attn_delta_kernel(const uint16_t* __restrict__ dO, int64_t lddo,
                                                         const uint16_t* __restrict__ O, int64_t ldo, int rows, int heads,
                                                         float* __restrict__ delta, int Lq, int Lq_stat, int64_t stat_stride) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pair = gid >> 3;  // (row, head)
  const int sub = (int)(gid & 7);
  const bool ok = pair < (int64_t)rows * heads;
  float s = 0.f;
  int row = 0, h = 0;
  if (ok) {
    row = (int)(pair / heads); h = (int)(pair % heads);
    const uint4 a = *reinterpret_cast<const uint4*>(dO + (int64_t)row * lddo + h * 64 + sub * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(O + (int64_t)row * ldo + h * 64 + sub * 8);
    const float2 a0 = unpack2<DT>(a.x), a1 = unpack2<DT>(a.y), a2 = unpack2<DT>(a.z), a3 = unpack2<DT>(a.w);
    const float2 b0 = unpack2<DT>(b.x), b1 = unpack2<DT>(b.y), b2 = unpack2<DT>(b.z), b3 = unpack2<DT>(b.w);
    s = (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y) +
        (a3.x * b3.x + a3.y * b3.y);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) delta[(int64_t)h * stat_stride + (int64_t)(row / Lq) * Lq_stat + row % Lq] = s;  // head-major, like the LSE
}

// patch gather: one block per token row, threads over k (coalesced along v within a patch row)
template <int DT>
__global__ void __launch_bounds__(128) patch_gather_kernel(const float* __restrict__ img, int C, int Himg, int Wimg,
                                                           int P, int ws, int G, uint16_t* __restrict__ out, int64_t ldo,
                                                           int Kpad) {
  const int t = blockIdx.x;  // window-major token index over the whole batch
  const int T = G * G;
  const int b = t / T, tt = t % T;
  const int nwx = G / ws;
  const int win = tt / (ws * ws), in = tt % (ws * ws);
  const int pi = (win / nwx) * ws + in / ws, pj = (win % nwx) * ws + in % ws;
  const int K = C * P * P;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < K) {
      const int c = k / (P * P), u = (k / P) % P, w = k % P;
      v = img[(((int64_t)b * C + c) * Himg + (pi * P + u)) * Wimg + pj * P + w];
    }
    if constexpr (DT == 0) reinterpret_cast<__half*>(out)[(int64_t)t * ldo + k] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(out)[(int64_t)t * ldo + k] = __float2bfloat16_rn(v);
  }
}

// window-major tokens [B][T][D] -> NCHW [B][D][G][G]: 32x32 smem transpose (token <-> channel)
__global__ void __launch_bounds__(256) tokens_to_nchw_kernel(const float* __restrict__ x, int G, int ws, int D,
                                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int T = G * G, nwx = G / ws;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;  // t0 indexes *spatial* (row-major) positions
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // read: for each spatial position s (32 of them), channel d0+tx  (coalesced over channels)
  for (int i = ty; i < 32; i += 8) {
    const int s = t0 + i;
    if (s < T) {
      const int pi = s / G, pj = s % G;
      const int tok = ((pi / ws) * nwx + pj / ws) * (ws * ws) + (pi % ws) * ws + (pj % ws);
      tile[i][tx] = x[((int64_t)b * T + tok) * D + d0 + tx];
    }
  }
  __syncthreads();
  // write: channel d0+i, spatial t0+tx (coalesced over spatial)
  for (int i = ty; i < 32; i += 8) {
    const int s = t0 + tx;
    if (s < T) out[((int64_t)b * D + d0 + i) * T + s] = tile[tx][i];
  }
}

template <int DT>
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ g, int G, int ws, int D,
                                                             float* __restrict__ dx, uint16_t* __restrict__ dx16,
                                                             int64_t ld16, const float* __restrict__ img_scale,
                                                             const float* __restrict__ gscale) {
  __shared__ float tile[32][33];
  const float gs = gscale != nullptr ? __ldg(gscale) : 1.f;
  const int T = G * G, nwx = G / ws;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int s = t0 + tx;
    if (s < T) tile[i][tx] = g[((int64_t)b * D + d0 + i) * T + s];  // tile[channel][spatial]
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int s = t0 + i;
    if (s < T) {
      const int pi = s / G, pj = s % G;
      const int tok = ((pi / ws) * nwx + pj / ws) * (ws * ws) + (pi % ws) * ws + (pj % ws);
      const float v = tile[tx][i] * gs;
      const int64_t r = (int64_t)b * T + tok;
      dx[r * D + d0 + tx] = v;
      if (dx16 != nullptr) {
        const float vs = img_scale != nullptr ? v * __ldg(img_scale + b) : v;
        if constexpr (DT == 0) reinterpret_cast<__half*>(dx16)[r * ld16 + d0 + tx] = __float2half_rn(vs);
        else reinterpret_cast<__nv_bfloat16*>(dx16)[r * ld16 + d0 + tx] = __float2bfloat16_rn(vs);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// LoRA operand packing
// ------------------------------------------------------------------------------------------
struct PackArgs {
  LoraSite s;
  uint16_t* down_T; uint16_t* w_ext; int64_t ldw; uint16_t* up_pack; uint16_t* wt_ext; int64_t ldwt;
};
template <int DT>
__global__ void __launch_bounds__(256) lora_pack_kernel(const PackArgs a) {
  const LoraSite& s = a.s;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // (1) down_T [rpad][in] and (4) wt_ext K-extension [in][rpad]
  for (int64_t i = tid; i < (int64_t)s.rpad * s.in; i += stride) {
    const int jj = (int)(i / s.in), k = (int)(i % s.in);
    const int ad = jj / s.r, j = jj % s.r;
    const float v = (ad < s.n) ? s.A[ad][(int64_t)k * s.r + j] : 0.f;
    const uint16_t h = to16<DT>(v);
    if (a.down_T) a.down_T[(int64_t)jj * s.in + k] = h;
    if (a.wt_ext) a.wt_ext[(int64_t)k * a.ldwt + s.out_total + jj] = h;
  }
  // (2) w_ext K-extension [out_total][rpad] and (3) up_pack [rpad][out_total]
  for (int64_t i = tid; i < (int64_t)s.rpad * s.out_total; i += stride) {
    const int jj = (int)(i / s.out_total), n = (int)(i % s.out_total);
    const int ad = jj / s.r, j = jj % s.r;
    float v = 0.f;
    if (ad < s.n && n >= s.out_off[ad] && n < s.out_off[ad] + s.out_len[ad])
      v = s.B[ad][(int64_t)j * s.out_len[ad] + (n - s.out_off[ad])];
    const uint16_t h = to16<DT>(v);
    if (a.w_ext) a.w_ext[(int64_t)n * a.ldw + s.in + jj] = h;
    if (a.up_pack) a.up_pack[(int64_t)jj * s.out_total + n] = h;
  }
}

struct UnpackArgs {
  LoraSite s;
  const float* dA_pack; const float* dB_pack;
  float* dA[3]; float* dB[3];
  const float* out_scale;
};
__global__ void __launch_bounds__(256) lora_unpack_kernel(const UnpackArgs a) {
  const LoraSite& s = a.s;
  const float sc = a.out_scale != nullptr ? __ldg(a.out_scale) : 1.f;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int ad = 0; ad < s.n; ++ad) {
    for (int64_t i = tid; i < (int64_t)s.in * s.r; i += stride) {
      const int k = (int)(i / s.r), j = (int)(i % s.r);
      a.dA[ad][i] = sc * a.dA_pack[(int64_t)k * s.rpad + ad * s.r + j];
    }
    for (int64_t i = tid; i < (int64_t)s.r * s.out_len[ad]; i += stride) {
      const int j = (int)(i / s.out_len[ad]), n = (int)(i % s.out_len[ad]);
      a.dB[ad][i] = sc * a.dB_pack[(int64_t)(ad * s.r + j) * s.out_total + s.out_off[ad] + n];
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256) lora_pack_all_kernel(const LoraSiteDesc* __restrict__ descs, const float* __restrict__ flat) {
  const LoraSiteDesc& s = descs[blockIdx.y];
  uint16_t* down_T = static_cast<uint16_t*>(s.down_T);
  uint16_t* w_ext = static_cast<uint16_t*>(s.w_ext);
  uint16_t* up_pack = static_cast<uint16_t*>(s.up_pack);
  uint16_t* wt_ext = static_cast<uint16_t*>(s.wt_ext);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < (int64_t)s.rpad * s.in; i += stride) {
    const int jj = (int)(i / s.in), k = (int)(i % s.in);
    const int ad = jj / s.r, j = jj % s.r;
    const float v = (ad < s.n) ? flat[s.a_off[ad] + (int64_t)k * s.r + j] : 0.f;
    const uint16_t h = to16<DT>(v);
    down_T[(int64_t)jj * s.in + k] = h;
    wt_ext[(int64_t)k * s.ldwt + s.out_total + jj] = h;
  }
  for (int64_t i = tid; i < (int64_t)s.rpad * s.out_total; i += stride) {
    const int jj = (int)(i / s.out_total), n = (int)(i % s.out_total);
    const int ad = jj / s.r, j = jj % s.r;
    float v = 0.f;
    if (ad < s.n && n >= s.out_off[ad] && n < s.out_off[ad] + s.out_len[ad])
      v = flat[s.b_off[ad] + (int64_t)j * s.out_len[ad] + (n - s.out_off[ad])];
    const uint16_t h = to16<DT>(v);
    w_ext[(int64_t)n * s.ldw + s.in + jj] = h;
    up_pack[(int64_t)jj * s.out_total + n] = h;
  }
}

__global__ void __launch_bounds__(256) lora_unpack_all_kernel(const LoraSiteDesc* __restrict__ descs, float* __restrict__ grad,
                                                              const float* __restrict__ out_scale) {
  const LoraSiteDesc& s = descs[blockIdx.y];
  const float sc = out_scale != nullptr ? __ldg(out_scale) : 1.f;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int ad = 0; ad < s.n; ++ad) {
    float* dA = grad + s.a_off[ad];
    float* dB = grad + s.b_off[ad];
    for (int64_t i = tid; i < (int64_t)s.in * s.r; i += stride) {
      const int k = (int)(i / s.r), j = (int)(i % s.r);
      dA[i] = sc * s.dA_pack[(int64_t)k * s.rpad + ad * s.r + j];
    }
    for (int64_t i = tid; i < (int64_t)s.r * s.out_len[ad]; i += stride) {
      const int j = (int)(i / s.out_len[ad]), n = (int)(i % s.out_len[ad]);
      dB[i] = sc * s.dB_pack[(int64_t)(ad * s.r + j) * s.out_total + s.out_off[ad] + n];
    }
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                    float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i];
    pi *= (1.f - lr * wd);  // decoupled weight decay (torch.optim.AdamW)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}


template <int VPT>
__global__ void __launch_bounds__(256) ln_fwd_f32_kernel(const float* x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, int rows, float* y) {
  constexpr int D = 128 * VPT;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * D);
  float4 v[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    v[i] = xr[lane + i * 32];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rs = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4* yr = reinterpret_cast<float4*>(y + (int64_t)row * D);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const float4 g = __ldg(g4 + lane + i * 32), b = __ldg(b4 + lane + i * 32);
    yr[lane + i * 32] = make_float4((v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y,
                                    (v[i].z - mu) * rs * g.z + b.z, (v[i].w - mu) * rs * g.w + b.w);
  }
}

// frozen weight cast (+ transpose through a 32x33 smem tile)
template <int DT>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ W, int N, int K, uint16_t* dst,
                                                          int64_t ld, int transpose) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + tx;
    tile[i][tx] = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
  }
  __syncthreads();
  if (!transpose) {
    for (int i = ty; i < 32; i += 8) {
      const int n = n0 + i, k = k0 + tx;
      if (n < N && k < K) dst[(int64_t)n * ld + k] = to16<DT>(tile[i][tx]);
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const int k = k0 + i, n = n0 + tx;
      if (n < N && k < K) dst[(int64_t)k * ld + n] = to16<DT>(tile[tx][i]);
    }
  }
}

__global__ void __launch_bounds__(256) build_pos_table_kernel(const float* __restrict__ pos, int side, int G, int ws,
                                                              int D, float* __restrict__ out) {
  const int t = blockIdx.x;
  const int nwx = G / ws;
  const int win = t / (ws * ws), in = t % (ws * ws);
  const int pi = (win / nwx) * ws + in / ws, pj = (win % nwx) * ws + in % ws;
  const float* src = pos + (int64_t)(1 + (pi % side) * side + (pj % side)) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[(int64_t)t * D + d] = src[d];
}


template <int DT>
__global__ void __launch_bounds__(256) dropout_rows16_kernel(const uint16_t* __restrict__ x, int64_t ldx, int rows, int cols,
                                                             uint16_t* __restrict__ out, int64_t ldo, float inv_keep,
                                                             uint32_t seed, uint32_t thr, const uint32_t* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed += __ldg(seed_dev);     // per-step seed kept in device memory (CUDA-graph replays)
  const int c8 = cols / 8;
  const int64_t n = (int64_t)rows * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / c8), c0 = (int)(i % c8) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(x + (int64_t)row * ldx + c0);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f = unpack2<DT>(w[q]);
      f.x = dropout_keep(seed, row, c0 + 2 * q, cols, thr) ? f.x * inv_keep : 0.f;
      f.y = dropout_keep(seed, row, c0 + 2 * q + 1, cols, thr) ? f.y * inv_keep : 0.f;
      o[q] = pack2<DT>(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(out + (int64_t)row * ldo + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

template <typename F>
int dispatch_vpt(int D, F&& f) {
  // one warp per row, VPT float4 per lane: any width that is a multiple of 128 up to 2048 (SAM3: 1024; ViT-B 768, ViT-H 1280)
  switch (D) {
    case 128: return f(std::integral_constant<int, 1>{});
    case 256: return f(std::integral_constant<int, 2>{});
    case 384: return f(std::integral_constant<int, 3>{});
    case 512: return f(std::integral_constant<int, 4>{});
    case 640: return f(std::integral_constant<int, 5>{});
    case 768: return f(std::integral_constant<int, 6>{});
    case 896: return f(std::integral_constant<int, 7>{});
    case 1024: return f(std::integral_constant<int, 8>{});
    case 1152: return f(std::integral_constant<int, 9>{});
    case 1280: return f(std::integral_constant<int, 10>{});
    case 1408: return f(std::integral_constant<int, 11>{});
    case 1536: return f(std::integral_constant<int, 12>{});
    case 1792: return f(std::integral_constant<int, 14>{});
    case 2048: return f(std::integral_constant<int, 16>{});
    default: return fail(-1, "layernorm: D=%d not supported (multiples of 128 up to 1536, 1792, 2048)", D);
  }
}

}  // namespace

int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int rows, int D, void* y16,
                  int64_t ldy, int dtype, float* mean, float* rstd, cudaStream_t s) {
  SAM3B_REQUIRE(ldy % 4 == 0, "layernorm: ldy %% 4 != 0");
  const int blocks = (rows + 7) / 8;
  return dispatch_vpt(D, [&](auto vpt) -> int {
    constexpr int V = decltype(vpt)::value;
    if (dtype == 0) SAM3B_CHECK_CUDA(launch_pdl(ln_fwd_kernel<V, 0>, dim3(blocks), dim3(256), 0, s, x, gamma, beta, eps, rows, (uint16_t*)y16, ldy, mean, rstd));
    else SAM3B_CHECK_CUDA(launch_pdl(ln_fwd_kernel<V, 1>, dim3(blocks), dim3(256), 0, s, x, gamma, beta, eps, rows, (uint16_t*)y16, ldy, mean, rstd));
    SAM3B_LAUNCHED();
    return 0;
  });
}

int layernorm_bwd(const void* dy16, int64_t lddy, const float* x, const float* mean, const float* rstd,
                  const float* gamma, const float* dres, int rows, int D, float* dx, void* dx16, int64_t lddx16,
                  int dtype, cudaStream_t s, const float* row_scale, int rows_per_scale) {
  SAM3B_REQUIRE(lddy % 4 == 0 && (dx16 == nullptr || lddx16 % 4 == 0), "layernorm bwd: ld %% 4 != 0");
  const int blocks = (rows + 7) / 8;
  return dispatch_vpt(D, [&](auto vpt) -> int {
    constexpr int V = decltype(vpt)::value;
    if (dtype == 0)
      SAM3B_CHECK_CUDA(launch_pdl(ln_bwd_kernel<V, 0>, dim3(blocks), dim3(256), 0, s, (const uint16_t*)dy16, lddy, x, mean, rstd, gamma, dres, rows, dx, (uint16_t*)dx16, lddx16, row_scale, rows_per_scale));
    else
      SAM3B_CHECK_CUDA(launch_pdl(ln_bwd_kernel<V, 1>, dim3(blocks), dim3(256), 0, s, (const uint16_t*)dy16, lddy, x, mean, rstd, gamma, dres, rows, dx, (uint16_t*)dx16, lddx16, row_scale, rows_per_scale));
    SAM3B_LAUNCHED();
    return 0;
  });
}


int layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float eps, int rows, int D, float* y32,
                      cudaStream_t s) {
  const int blocks = (rows + 7) / 8;
  return dispatch_vpt(D, [&](auto vpt) -> int {
    constexpr int V = decltype(vpt)::value;
    ln_fwd_f32_kernel<V><<<blocks, 256, 0, s>>>(x, gamma, beta, eps, rows, y32);
    SAM3B_LAUNCHED();
    return 0;
  });
}

int pack_weight(const float* W, int N, int K, void* dst16, int64_t ld, int transpose, int dtype, cudaStream_t s) {
  dim3 grid((K + 31) / 32, (N + 31) / 32);
  if (dtype == 0) pack_weight_kernel<0><<<grid, 256, 0, s>>>(W, N, K, (uint16_t*)dst16, ld, transpose);
  else pack_weight_kernel<1><<<grid, 256, 0, s>>>(W, N, K, (uint16_t*)dst16, ld, transpose);
  SAM3B_LAUNCHED();
  return 0;
}

int build_pos_table(const float* pos_embed, int side, int G, int ws, int D, float* out, cudaStream_t s) {
  build_pos_table_kernel<<<G * G, 256, 0, s>>>(pos_embed, side, G, ws, D, out);
  SAM3B_LAUNCHED();
  return 0;
}

int dropout_rows16(const void* x16, int64_t ldx, int rows, int cols, void* out16, int64_t ldo, float p, uint32_t seed,
                   int dtype, cudaStream_t s, const uint32_t* seed_dev) {
  SAM3B_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "dropout_rows16: cols / ld must be multiples of 8");
  SAM3B_REQUIRE(p >= 0.f && p < 1.f, "dropout_rows16: p=%f outside [0,1)", p);
  const int64_t n = (int64_t)rows * (cols / 8);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  const float inv_keep = 1.f / (1.f - p);
  const uint32_t thr = dropout_threshold(p);
  if (dtype == 0) dropout_rows16_kernel<0><<<blocks, 256, 0, s>>>((const uint16_t*)x16, ldx, rows, cols, (uint16_t*)out16, ldo, inv_keep, seed, thr, seed_dev);
  else dropout_rows16_kernel<1><<<blocks, 256, 0, s>>>((const uint16_t*)x16, ldx, rows, cols, (uint16_t*)out16, ldo, inv_keep, seed, thr, seed_dev);
  SAM3B_LAUNCHED();
  return 0;
}

int cast_rows_16(const float* x, int rows, int D, void* y16, int64_t ldy, int dtype, cudaStream_t s, const float* scale) {
  SAM3B_REQUIRE(D % 4 == 0 && ldy % 4 == 0, "cast: D and ldy must be multiples of 4");
  const int64_t n4 = (int64_t)rows * (D / 4);
  const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)num_sms() * 16);
  if (dtype == 0) cast_rows_kernel<0><<<blocks, 256, 0, s>>>(x, rows, D, (uint16_t*)y16, ldy, scale);
  else cast_rows_kernel<1><<<blocks, 256, 0, s>>>(x, rows, D, (uint16_t*)y16, ldy, scale);
  SAM3B_LAUNCHED();
  return 0;
}

int attn_delta(const void* dO, int64_t lddo, const void* O, int64_t ldo, int rows, int heads, int dtype, float* delta,
               cudaStream_t s, int Lq) {
  if (Lq <= 0) Lq = rows;
  SAM3B_REQUIRE(rows % Lq == 0, "attn_delta: rows %d not a multiple of Lq %d", rows, Lq);
  const int Lq_stat = (Lq + 63) / 64 * 64;
  const int64_t stat_stride = (int64_t)(rows / Lq) * Lq_stat;
  SAM3B_REQUIRE(lddo % 8 == 0 && ldo % 8 == 0, "attn_delta: ld %% 8 != 0");
  const int64_t threads = (int64_t)rows * heads * 8;
  const int blocks = (int)((threads + 255) / 256);
  if (dtype == 0) attn_delta_kernel<0><<<blocks, 256, 0, s>>>((const uint16_t*)dO, lddo, (const uint16_t*)O, ldo, rows, heads, delta, Lq, Lq_stat, stat_stride);
  else attn_delta_kernel<1><<<blocks, 256, 0, s>>>((const uint16_t*)dO, lddo, (const uint16_t*)O, ldo, rows, heads, delta, Lq, Lq_stat, stat_stride);
  SAM3B_LAUNCHED();
  return 0;
}

int patch_gather(const float* img, int B, int C, int Himg, int Wimg, int P, int ws, void* out16, int64_t ldo, int Kpad,
                 int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(Himg == Wimg && Himg % P == 0, "patch_gather: square image with side %% patch == 0 expected");
  const int G = Himg / P;
  SAM3B_REQUIRE(G % ws == 0, "patch_gather: grid %d not a multiple of window %d", G, ws);
  SAM3B_REQUIRE(Kpad >= C * P * P && Kpad <= ldo, "patch_gather: Kpad");
  const int tokens = B * G * G;
  if (dtype == 0) patch_gather_kernel<0><<<tokens, 128, 0, s>>>(img, C, Himg, Wimg, P, ws, G, (uint16_t*)out16, ldo, Kpad);
  else patch_gather_kernel<1><<<tokens, 128, 0, s>>>(img, C, Himg, Wimg, P, ws, G, (uint16_t*)out16, ldo, Kpad);
  SAM3B_LAUNCHED();
  return 0;
}

int tokens_to_nchw(const float* x, int B, int G, int ws, int D, float* out, cudaStream_t s) {
  SAM3B_REQUIRE(D % 32 == 0 && G % ws == 0, "tokens_to_nchw: D %% 32, G %% ws");
  dim3 grid((G * G + 31) / 32, D / 32, B);
  tokens_to_nchw_kernel<<<grid, 256, 0, s>>>(x, G, ws, D, out);
  SAM3B_LAUNCHED();
  return 0;
}

int nchw_to_tokens(const float* g, int B, int G, int ws, int D, float* dx, void* dx16, int64_t ld16, int dtype,
                   cudaStream_t s, const float* img_scale, const float* gscale) {
  SAM3B_REQUIRE(D % 32 == 0 && G % ws == 0, "nchw_to_tokens: D %% 32, G %% ws");
  dim3 grid((G * G + 31) / 32, D / 32, B);
  if (dtype == 0) nchw_to_tokens_kernel<0><<<grid, 256, 0, s>>>(g, G, ws, D, dx, (uint16_t*)dx16, ld16, img_scale, gscale);
  else nchw_to_tokens_kernel<1><<<grid, 256, 0, s>>>(g, G, ws, D, dx, (uint16_t*)dx16, ld16, img_scale, gscale);
  SAM3B_LAUNCHED();
  return 0;
}

int lora_pack(const LoraSite& site, void* down_T, void* w_ext, int64_t ldw, void* up_pack, void* wt_ext, int64_t ldwt,
              int dtype, cudaStream_t s) {
  SAM3B_REQUIRE(site.n >= 1 && site.n <= 3 && site.n * site.r <= site.rpad, "lora_pack: %d adapters of rank %d exceed rpad %d", site.n, site.r, site.rpad);
  PackArgs a{site, (uint16_t*)down_T, (uint16_t*)w_ext, ldw, (uint16_t*)up_pack, (uint16_t*)wt_ext, ldwt};
  const int64_t work = (int64_t)site.rpad * std::max(site.in, site.out_total);
  const int blocks = (int)std::min<int64_t>((work + 255) / 256, 1024);
  if (dtype == 0) lora_pack_kernel<0><<<blocks, 256, 0, s>>>(a);
  else lora_pack_kernel<1><<<blocks, 256, 0, s>>>(a);
  SAM3B_LAUNCHED();
  return 0;
}

int lora_unpack_grads(const LoraSite& site, const float* dA_pack, const float* dB_pack, float* const dA[3],
                      float* const dB[3], cudaStream_t s, const float* out_scale) {
  UnpackArgs a{};
  a.s = site; a.dA_pack = dA_pack; a.dB_pack = dB_pack; a.out_scale = out_scale;
  for (int i = 0; i < 3; ++i) { a.dA[i] = dA[i]; a.dB[i] = dB[i]; }
  const int64_t work = (int64_t)site.r * std::max(site.in, site.out_total);
  const int blocks = (int)std::min<int64_t>((work + 255) / 256, 512);
  lora_unpack_kernel<<<blocks, 256, 0, s>>>(a);
  SAM3B_LAUNCHED();
  return 0;
}

int lora_pack_all(const LoraSiteDesc* dev_descs, int n_sites, int max_work, const float* flat, int dtype, cudaStream_t s) {
  if (n_sites <= 0) return 0;
  SAM3B_REQUIRE(dev_descs && flat && n_sites <= 65535, "lora_pack_all: bad arguments");
  const dim3 grid(std::max(1, std::min((max_work + 255) / 256, 64)), n_sites);
  if (dtype == 0) lora_pack_all_kernel<0><<<grid, 256, 0, s>>>(dev_descs, flat);
  else lora_pack_all_kernel<1><<<grid, 256, 0, s>>>(dev_descs, flat);
  SAM3B_LAUNCHED();
  return 0;
}

int lora_unpack_all(const LoraSiteDesc* dev_descs, int n_sites, int max_work, float* grad_flat, const float* out_scale,
                    cudaStream_t s) {
  if (n_sites <= 0) return 0;
  SAM3B_REQUIRE(dev_descs && grad_flat && n_sites <= 65535, "lora_unpack_all: bad arguments");
  const dim3 grid(std::max(1, std::min((max_work + 255) / 256, 64)), n_sites);
  lora_unpack_all_kernel<<<grid, 256, 0, s>>>(dev_descs, grad_flat, out_scale);
  SAM3B_LAUNCHED();
  return 0;
}

int adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, cudaStream_t s) {
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8);
  adamw_kernel<<<blocks, 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale);
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace sam3b
