#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {
// loss (elementwise, may be null) and/or sum (scalar, zeroed by the call, may be null)
int focal_loss_fwd(const float* x, const float* y, int64_t n, float alpha, float gamma, float* loss, float* sum, cudaStream_t s);
// dx = dL/dx * gscale * (g ? g[i] : 1)
int focal_loss_bwd(const float* x, const float* y, int64_t n, float alpha, float gamma, const float* g, float gscale, float* dx,
                   cudaStream_t s);
}  // namespace sam3b
