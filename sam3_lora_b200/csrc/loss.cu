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

}  // namespace sam3b
