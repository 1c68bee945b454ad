// Micro-benchmark: issue rate of tcgen05.mma kind::f16, M=128, K=16, for several N and operand sources.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sam3_lora_b200/csrc tools/micro/mma_rate.cu -o gpurun_out/mma_rate
// Prints cycles per MMA instruction (one issuing thread per SM, all SMs busy) for
//   SS N=64 / TS N=64 / TS N=128 / SS N=256, with the accumulator alternating between two TMEM regions.
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace sam3b;

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(unsigned long long* out, int rounds) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, 0, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    const unsigned long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tb + ((r % NACC) * N) % 256;
        if (TS) umma_f16_ts(d, tb + 448 + k * 8, make_desc_kmajor(b_addr + k * 32), idesc, 1);
        else    umma_f16_ss(d, make_desc_kmajor(a_addr + k * 32), make_desc_kmajor(b_addr + k * 32), idesc, 1);
      }
    }
    const unsigned long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const unsigned long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

template <int N, bool TS, int NACC>
void run(const char* name, unsigned long long* d_out) {
  const int rounds = 2000;
  const int smem = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<N, TS, NACC><<<148, 128, smem>>>(d_out, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n = rounds * 4.0;
  printf("{\"case\": \"%s\", \"err\": \"%s\", \"issue_clk_per_mma\": %.1f, \"complete_clk_per_mma\": %.1f, \"ideal_clk\": %.1f}\n", name,
         cudaGetErrorString(e), h[0] / n, h[1] / n, N / 2.0);
}

// Same TS N=64 stream, with the other warps of the CTA generating traffic: MODE 1 = tcgen05.ld x32 loops,
// 2 = tcgen05.st x32 loops, 3 = MUFU/FMA loops, 4 = bulk (TMA) copies global -> smem, 5 = ld + st + MUFU (a softmax-like mix)
template <int MODE>
__global__ void __launch_bounds__(576, 1) mma_contend_kernel(unsigned long long* out, int rounds, const float* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, tbar[2];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); stop = 0; fence_barrier_init(); }
  if (warp == 17) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (16384 + 64 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 17) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, 64, 0, 0, 0);
      const uint32_t b_addr = smem_u32(smem + 16384);
      const unsigned long long t0 = clock64();
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tb + 256 + (r & 1) * 64, tb + 448 + k * 8, make_desc_kmajor(b_addr + k * 32), idesc, 1);
      }
      const unsigned long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0, 1);
      const unsigned long long t2 = clock64();
      stop = 1;
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (warp == 16) {
    if (MODE == 4 && elect_one()) {
      int ph[2] = {0, 0};
      for (int i = 0; !stop; ++i) {
        const int b = i & 1;
        mbar_arrive_expect_tx(&tbar[b], 16384);
        bulk_load_1d(smem + 32768 + b * 16384, gsrc + (size_t)((i * 148 + blockIdx.x) % 4096) * 4096, 16384, &tbar[b]);
        mbar_wait(&tbar[b], ph[b], 2);
        ph[b] ^= 1;
      }
    }
  } else {
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t col = ((warp >> 2) & 3) * 64;   // columns [0,256): disjoint from the MMA's A / D columns
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    float acc = threadIdx.x * 1e-3f;
    while (!stop) {
      if (MODE == 1 || MODE == 5) { tmem_ld_x32(tb + lane_off + col, v); tmem_ld_wait(); }
      if (MODE == 3 || MODE == 5) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { acc = ex2_approx(acc * 0.5f - __uint_as_float(v[i] & 0x3f800000u)); v[i] ^= __float_as_uint(acc) & 1u; }
      }
      if (MODE == 2 || MODE == 5) { tmem_st_x32(tb + lane_off + col + 32, v); tmem_st_wait(); }
    }
    if (acc == 123.456f) out[7] = v[3];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tb, 512);
}

template <int MODE>
void run_contend(const char* name, unsigned long long* d_out, const float* gsrc) {
  const int rounds = 4000;
  const int smem = 65536 + 1024;
  cudaFuncSetAttribute(mma_contend_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_contend_kernel<MODE><<<148, 576, smem>>>(d_out, rounds, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n = rounds * 4.0;
  printf("{\"case\": \"TS N=64 + %s\", \"err\": \"%s\", \"issue_clk_per_mma\": %.1f, \"complete_clk_per_mma\": %.1f, \"ideal_clk\": 32.0}\n", name,
         cudaGetErrorString(e), h[0] / n, h[1] / n);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 64);
  run<64, false, 1>("SS N=64 one accumulator", d_out);
  run<64, false, 2>("SS N=64 two accumulators", d_out);
  run<64, true, 1>("TS N=64 one accumulator", d_out);
  run<64, true, 2>("TS N=64 two accumulators", d_out);
  run<128, true, 1>("TS N=128 one accumulator", d_out);
  run<128, false, 1>("SS N=128 one accumulator", d_out);
  run<256, false, 1>("SS N=256 one accumulator", d_out);
  run<256, true, 1>("TS N=256 one accumulator", d_out);
  float* gsrc;
  cudaMalloc(&gsrc, (size_t)4096 * 4096 * 4 + 65536);
  cudaMemset(gsrc, 0, (size_t)4096 * 4096 * 4 + 65536);
  run_contend<0>("idle warps", d_out, gsrc);
  run_contend<1>("16 warps tcgen05.ld", d_out, gsrc);
  run_contend<2>("16 warps tcgen05.st", d_out, gsrc);
  run_contend<3>("16 warps MUFU", d_out, gsrc);
  run_contend<4>("TMA bulk copies into smem", d_out, gsrc);
  run_contend<5>("16 warps ld+MUFU+st", d_out, gsrc);
  return 0;
}
