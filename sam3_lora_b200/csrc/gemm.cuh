// Internal GEMM interface (host side). The C ABI wrapper in capi.cpp and the block
// schedules in engine.cpp both go through gemm_launch().
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

enum GemmEpilogue : int {
  EPI_STORE16 = 0,       // C16 = alpha*acc (+bias)
  EPI_QKV_ROPE = 1,      // C16 = rope(acc + bias) on columns < rope_cols, plain on the rest
  EPI_RESIDUAL_F32 = 2,  // C32 = acc + bias + residual[row % res_row_mod]
  EPI_GELU = 3,          // C16 = h = acc + bias ; C2_16 = gelu_erf(h)
  EPI_DGELU = 4,         // C16 = acc * gelu_erf'(aux16)
  EPI_ATOMIC_F32 = 5,    // C32 (+)= alpha*acc with red.global.add (split-K), optional transpose
  EPI_STORE32 = 6,       // C32 = [row_scale[row/rows_per_scale] *] (alpha*acc (+bias))
  EPI_ADDMASK16 = 7,     // C16 += keep(row,col)/(1-p) * alpha*acc [* gelu_erf'(aux16)]   (adapter dgrad under dropout)
  EPI_STORE16_DELTA = 8, // C16 = acc ; delta[col/64][row'] += sum_c C16[row][c] * aux16[row][c]  (proj dgrad: dO and the
                         // flash-attention backward's delta = rowsum(dO * O) per head in one pass; engine-internal)
  EPI_COUNT
};

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr; int64_t lda = 0; int a_mn = 0;  // a_mn=0: A[M][K] ; 1: A stored as [K][M]
  const void* B = nullptr; int64_t ldb = 0; int b_mn = 0;  // b_mn=0: B[N][K] ; 1: B stored as [K][N]
  int dtype = 0;                                           // operand format: 0 fp16, 1 bf16
  int epilogue = EPI_STORE16;
  void* C = nullptr; int64_t ldc = 0;
  void* C2 = nullptr; int64_t ldc2 = 0;
  const float* bias = nullptr;
  const float* residual = nullptr; int64_t ldres = 0; int res_row_mod = 0;
  float drop_p = 0.f; uint32_t drop_seed = 0;   // EPI_ADDMASK16: mask = rng.cuh dropout_keep(seed, row, col, N)
  const uint32_t* drop_seed_dev = nullptr;      // optional device word added to drop_seed (per-step seed under CUDA graphs)
  const float* row_scale = nullptr; int rows_per_scale = 1;  // EPI_RESIDUAL_F32: out = res + row_scale[row / rows_per_scale] * (acc + bias)  (DropPath)
  const void* aux = nullptr; int64_t ldaux = 0;
  const float* rope = nullptr; int rope_period = 1; int rope_cols = 0;  // rope: [period][32] (cos,sin) pairs
  float alpha = 1.f;
  int splitk = 1;
  int c_trans = 0;      // EPI_ATOMIC_F32 only: write C[col*ldc + row]
  int bn = 0;           // 0 = choose; else 64 or 256
  int dbg_lbo = 0, dbg_sbo = 0;  // bring-up overrides for the MN-major descriptors (bytes); 0 = default
  float* delta = nullptr; int delta_Lq = 0, delta_Lq_stat = 0; int64_t delta_stride = 0;  // EPI_STORE16_DELTA: head-major
                        // statistics layout of the attention kernels: delta[h * stride + (row / Lq) * Lq_stat + row % Lq], zeroed by the caller
  int max_ctas = 0;     // 0 = number of SMs
  int cta_pair = 0;     // 0 = default policy, 1 = single-CTA tiles, 2 = CTA-pair (cta_group::2) 256x256 tiles
};

int gemm_launch(const GemmArgs& a, cudaStream_t stream);

// Implicit-GEMM 3x3 convolution, stride 1, zero padding 1, on channels-last 16-bit activations (nn.Conv2d(C, Cout, 3,
// padding=1): necks.py:84-92, maskformer_segmentation.py:187).  No im2col buffer: the A tile of filter tap (ky, kx) is ONE
// 4-D TMA box of the input shifted by (ky-1, kx-1), with out-of-range pixels zero-filled by the TMA unit.
//   x16 [B][H][W][C]   w9 [Cout][9*C] with k = (ky*3 + kx)*C + c   out [B*H*W][ldc] 16-bit or fp32 (+ bias [Cout])
// Needs C % 64 == 0, W % 8 == 0, Cout % 8 == 0 (callers fall back to im2col3x3 + gemm_launch otherwise).
struct ConvArgs {
  int B = 0, H = 0, W = 0, C = 0, Cout = 0;
  const void* x16 = nullptr;
  const void* w9 = nullptr;
  const float* bias = nullptr;
  void* out = nullptr; int64_t ldc = 0;
  int out_f32 = 0;
  int dtype = 0;
};
bool conv3x3_supported(int H, int W, int C, int Cout);
int conv3x3_launch(const ConvArgs& a, cudaStream_t stream);

}  // namespace sam3b
