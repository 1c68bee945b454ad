/*
 * sam3b.h — C ABI of libsam3b.so, the B200 (sm_100a) hot path behind the sam3_lora
 * LoRA-training surface.
 *
 * The reference (Sompote/sam3_lora) is pure Python/PyTorch: it has no FFI of its own
 * (SURVEY.md §0.1, §8b).  The "binding a maintainer would add" is therefore a ctypes stub
 *  that replaces the PyTorch-eager math of these reference call sites:
 *
 *   sam3b_gemm            nn.Linear + LoRALinear.forward        lora_layers.py:49-55,87-91
 *                         Attention.qkv / .proj                 sam3/model/vitdet.py:480,513
 *                         timm Mlp fc1 / GELU / fc2             sam3/model/vitdet.py:585-590,611
 *   sam3b_layernorm_*     nn.LayerNorm(eps=1e-5)                sam3/model/vitdet.py:566,584,719,833
 *   sam3b_attention_*     apply_rotary_enc + F.scaled_dot_product_attention
 *                                                               sam3/model/vitdet.py:68-90,485,502
 *   sam3b_patch_gather    PatchEmbed conv (k=s=14) im2col side  sam3/model/vitdet.py:323-336
 *   sam3b_vit_*           ViT.forward / Block.forward (+ autograd backward)
 *                                                               sam3/model/vitdet.py:597-613,813-859
 *
 * Conventions: every function returns 0 on success or a negative code; the message for the
 * last failure on the calling thread is returned by sam3b_last_error().  All pointers are
 * device pointers unless stated otherwise; the library never takes ownership and never
 * allocates device memory (workspaces are passed in).  `stream` is a cudaStream_t passed
 * as void*.  No exceptions cross the boundary; the library keeps no global mutable state
 * besides per-kernel attribute caches.
 */
#ifndef SAM3B_H_
#define SAM3B_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAM3B_ABI_VERSION 1

/* operand formats of the tensor-core path (fp32 accumulate always) */
#define SAM3B_F16 0
#define SAM3B_BF16 1

/* GEMM epilogues */
#define SAM3B_EPI_STORE16 0      /* C16 = alpha*acc (+bias) */
#define SAM3B_EPI_QKV_ROPE 1     /* C16 = rope(acc+bias) on columns < rope_cols */
#define SAM3B_EPI_RESIDUAL_F32 2 /* C32 = acc + bias + residual[row % res_row_mod] */
#define SAM3B_EPI_GELU 3         /* C16 = h = acc+bias ; C2_16 = gelu_erf(h) */
#define SAM3B_EPI_DGELU 4        /* C16 = acc * gelu_erf'(aux16) */
#define SAM3B_EPI_ATOMIC_F32 5   /* C32 += alpha*acc (split-K, red.global.add) */
#define SAM3B_EPI_STORE32 6      /* C32 = alpha*acc (+bias) */

const char* sam3b_last_error(void);
int sam3b_abi_version(void);

/* C[M][N] = epilogue(alpha * A[M][K] . B[N][K]^T), 16-bit operands, fp32 accumulation in TMEM. */
typedef struct sam3b_gemm_desc {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_mn; /* a_mn=0: A is [M][K]; 1: A is stored [K][M] */
  const void* B; int64_t ldb; int32_t b_mn; /* b_mn=0: B is [N][K]; 1: B is stored [K][N] */
  int32_t dtype;                            /* SAM3B_F16 | SAM3B_BF16 */
  int32_t epilogue;                         /* SAM3B_EPI_* */
  void* C; int64_t ldc;
  void* C2; int64_t ldc2;
  const float* bias;                        /* [N] or NULL */
  const float* residual; int64_t ldres; int32_t res_row_mod;
  const void* aux; int64_t ldaux;
  const float* rope; int32_t rope_period; int32_t rope_cols; /* rope: [period][32][2] (cos,sin) */
  float alpha;
  int32_t splitk;
  int32_t c_trans;
  int32_t bn;                               /* 0 = auto, 64, 256 */
  int32_t dbg_lbo, dbg_sbo;                 /* bring-up only; 0 = default */
  int32_t max_ctas;                         /* 0 = one CTA per SM */
} sam3b_gemm_desc;

int sam3b_gemm(const sam3b_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAM3B_H_ */
