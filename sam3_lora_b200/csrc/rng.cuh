// Stateless dropout mask for the LoRA adapter branch (lora_layers.py:43,54: nn.Dropout on the adapter input).
// keep(row, col) is a pure function of (site seed, row, col) so forward, weight-gradient and data-gradient
// kernels regenerate the same mask instead of storing it; the oracle restates the same hash in numpy.
#pragma once
#include <cstdint>

namespace sam3b {

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// true = element is kept.  `cols` is the row pitch of the logical [rows][cols] activation.
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t row, uint32_t col, uint32_t cols, uint32_t thr) {
  return lowbias32(seed ^ lowbias32(row * cols + col)) >= thr;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
__host__ __device__ __forceinline__ uint32_t site_seed(uint32_t seed, int block, int site) {
  return seed + 0x9E3779B9u * (uint32_t)(block * 4 + site + 1);
}

}  // namespace sam3b
