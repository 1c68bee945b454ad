// Internal interface of the attention kernels (host side).
//
// One kernel family serves two reference call-site families:
//   * ViT self-attention (vitdet.py:485-502): q|k|v column blocks of one buffer, Lq == Lk, head_dim 64, no mask.
//   * nn.MultiheadAttention sites of the DETR encoder/decoder/seg head (model_misc.py:31-34; encoder.py:139-201,
//     decoder.py:80-187, maskformer_segmentation.py:281-289): head_dim 32 stored zero-padded to 64 columns per head,
//     separate query / key-value buffers (cross attention), additive float attn_mask, boolean key_padding_mask and
//     dropout on the attention probabilities.  These extras compile into a second instantiation (GEN) so the ViT
//     instantiation carries none of their instructions.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

struct AttnArgs {
  // operands: 16-bit, head h of q at column q_col0 + 64*h of `q`, of k / v at k_col0 / v_col0 + 64*h of `kv`
  const void* q = nullptr; int64_t ldq = 0; int q_cols = 0; int q_col0 = 0;
  const void* kv = nullptr; int64_t ldkv = 0; int kv_cols = 0; int k_col0 = 0, v_col0 = 0;
  int nseg = 0;        // independent attention problems (windows / images / batch elements)
  int Lq = 0, Lk = 0;  // rows per segment in q and in kv (row = seg*L + i)
  int heads = 0;
  float scale = 0.125f;  // head_dim^-0.5 of the *logical* head dim
  int dtype = 0;         // 0 fp16, 1 bf16
  void* O = nullptr; int64_t ldo = 0; int o_col0 = 0;   // [nseg*Lq][>= o_col0 + 64*heads] 16-bit
  float* lse2 = nullptr;   // [heads][nseg*Lq_stat] log2-domain log-sum-exp; Lq_stat = Lq rounded up to 64
  // optional (any of them selects the GEN instantiation)
  const float* bias = nullptr;       // additive attn_mask, fp32 [nseg*heads][Lq][Lk] (natural-log units, like torch)
  const uint8_t* kpm = nullptr;      // key_padding_mask [nseg][Lk], non-zero = ignore key
  float drop_p = 0.f; uint32_t drop_seed = 0;  // dropout on the attention probabilities (training)
  // optional precomputed keep-bits of that same mask (attn_dropout_bits): bit (k & 31) of drop_bits[(bh*Lq + q)*pitch(Lk) + k/32]
  // and bit (q & 31) of drop_bitsT[(bh*Lk + k)*pitch(Lq) + q/32], bh = seg*heads + head, pitch = attn_bits_pitch.  The three kernels then read one word
  // per 32 scores instead of hashing every score three times (forward, dK/dV, dQ).  Needs Lq % 32 == 0 and Lk % 32 == 0.
  const uint32_t* drop_bits = nullptr; const uint32_t* drop_bitsT = nullptr;
  // backward only
  const void* dO = nullptr; int64_t lddo = 0; int do_col0 = 0;
  const float* delta = nullptr;      // [heads][nseg*Lq_stat] rowsum(dO * O)
  void* dq = nullptr; int64_t lddq = 0; int dq_col0 = 0;        // gradient w.r.t. q (same layout as q)
  void* dkv = nullptr; int64_t lddkv = 0; int dk_col0 = 0, dv_col0 = 0;
  const float* rope = nullptr; int rope_period = 1;             // optional inverse RoPE on dq, dk (ViT)
};

inline int attn_lq_stat(int Lq) { return (Lq + 63) / 64 * 64; }

int attn_fwd_launch(const AttnArgs& a, cudaStream_t stream);
// keep-bits of the attention-dropout mask for bh = 0 .. n_bh-1 in both orientations (see AttnArgs::drop_bits)
// row pitch (32-bit words) of the keep-bit arrays for a dimension of length L: L/32 rounded up to 8 words (32-byte rows)
int attn_bits_pitch(int L);
int attn_dropout_bits(int n_bh, int Lq, int Lk, float p, uint32_t seed, uint32_t* bits, uint32_t* bitsT, cudaStream_t stream);
int attn_bwd_launch(const AttnArgs& a, cudaStream_t stream);

}  // namespace sam3b
