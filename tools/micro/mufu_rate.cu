// Micro-benchmark: MUFU.EX2 throughput per SM sub-partition, fp32 vs packed f16x2 (two exponentials per instruction?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/micro/mufu_rate.cu -o tools/micro/mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) mufu_kernel(unsigned long long* out, float seed, int iters) {
  float xf[8];
  uint32_t xh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { xf[i] = seed * (threadIdx.x + i) * 1e-4f - 1.f; xh[i] = 0xB800B400u + threadIdx.x + i; }  // halves ~ -0.25 / -0.5
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(xf[i]));
      else if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(xh[i]));
      else asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(xh[i]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { xf[i] -= 1.5f; xh[i] ^= 0x80008000u; }   // keep values in range (one FADD / LOP per exp)
  }
  const unsigned long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += xf[i] + __uint_as_float(xh[i]);
  if (acc == 123.456f) out[7] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, unsigned long long* d) {
  const int iters = 2000;
  mufu_kernel<MODE><<<148, 512>>>(d, 1.f, iters);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  // 16 warps per SM = 4 per sub-partition, 8 instructions per iteration each
  const double per_warp_instr = (double)h / (iters * 8.0 * 4.0);
  printf("{\"case\": \"%s\", \"err\": \"%s\", \"clk_per_warp_instruction_per_subpartition\": %.2f}\n", name, cudaGetErrorString(e), per_warp_instr);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  run<0>("ex2.approx.ftz.f32", d);
  run<1>("ex2.approx.ftz.f16x2", d);
  run<2>("ex2.approx.ftz.bf16x2", d);
  return 0;
}
