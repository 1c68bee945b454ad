// Internal interface of the attention kernels (host side).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

struct AttnFwdArgs {
  const void* qkv = nullptr; int64_t ldqkv = 0;  // [tokens][>=3D] 16-bit: q | k | v column blocks of width D
  int tokens = 0;      // total rows
  int seg_len = 0;     // tokens per attention segment (576 window / 5184 image)
  int D = 0, heads = 0, head_dim = 64;
  int dtype = 0;       // 0 fp16, 1 bf16
  void* O = nullptr; int64_t ldo = 0;            // [tokens][>=D] 16-bit, head h at columns h*64
  float* lse2 = nullptr;                         // [heads][tokens] log2-domain log-sum-exp
};
int attn_fwd_launch(const AttnFwdArgs& a, cudaStream_t stream);

struct AttnBwdArgs {
  const void* qkv = nullptr; int64_t ldqkv = 0;  // forward q|k|v (rotated q,k)
  const void* dO = nullptr; int64_t lddo = 0;    // [tokens][>=D] 16-bit
  const float* lse2 = nullptr;                   // [heads][tokens]
  const float* delta = nullptr;                  // [heads][tokens] rowsum(dO * O)
  void* dqkv = nullptr; int64_t lddqkv = 0;      // [tokens][>=3D] 16-bit out: dq | dk | dv (un-rotated q,k grads)
  const float* rope = nullptr; int rope_period = 1;  // [period][32][2] (cos,sin); inverse rotation on dq, dk
  int tokens = 0, seg_len = 0, D = 0, heads = 0, head_dim = 64;
  int dtype = 0;
};
int attn_bwd_launch(const AttnBwdArgs& a, cudaStream_t stream);

}  // namespace sam3b
