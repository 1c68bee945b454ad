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

/* ---- LayerNorm (one warp per token row) -------------------------------------------------- */
/* y16[row][0..D) = LN(x[row]) * gamma + beta ; saves mean/rstd.   nn.LayerNorm(eps=1e-5), vitdet.py:566,584,719,833 */
int sam3b_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t D,
                        void* y16, int64_t ldy, int32_t dtype, float* mean, float* rstd, void* stream);
/* dx = dres + dLN(dy16) (autograd of the above + the residual skip); dx16 = 16-bit copy of dx (may be NULL) */
int sam3b_layernorm_bwd(const void* dy16, int64_t lddy, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* dres, int32_t rows, int32_t D, float* dx, void* dx16,
                        int64_t lddx16, int32_t dtype, void* stream);
int sam3b_cast_rows_16(const float* x, int32_t rows, int32_t D, void* y16, int64_t ldy, int32_t dtype, void* stream);

/* ---- attention (head_dim 64; tokens in window-major order so a window/image is a row run) --- */
/* apply_rotary_enc + F.scaled_dot_product_attention, vitdet.py:68-90,485,502 (RoPE itself is the
 * SAM3B_EPI_QKV_ROPE epilogue of the qkv GEMM). */
typedef struct sam3b_attn_desc {
  const void* qkv; int64_t ldqkv;   /* [tokens][>=3D] 16-bit: q | k | v, head h at columns h*64 of each block */
  int32_t tokens, seg_len, D, heads, head_dim, dtype;
  void* O; int64_t ldo;             /* fwd out / bwd in: [tokens][>=D] 16-bit */
  float* lse2;                      /* [tokens][heads], log2-domain log-sum-exp (fwd out / bwd in) */
  /* backward only */
  const void* dO; int64_t lddo;
  float* delta;                     /* [tokens][heads] scratch: rowsum(dO*O), written by the call */
  void* dqkv; int64_t lddqkv;       /* [tokens][>=3D] 16-bit out: gradients w.r.t. the un-rotated q | k | v */
  const float* rope; int32_t rope_period;
} sam3b_attn_desc;
int sam3b_attention_fwd(const sam3b_attn_desc* d, void* stream);
int sam3b_attention_bwd(const sam3b_attn_desc* d, void* stream);

/* ---- patch embed gather + layout ----------------------------------------------------------- */
/* PatchEmbed conv k=s=P as a GEMM (vitdet.py:323-336): gathers 16-bit rows [token][Kpad], k=(c*P+u)*P+v,
 * tokens in window-major order (window_partition, vitdet.py:93-115, folded into the layout). */
int sam3b_patch_gather(const float* img, int32_t B, int32_t C, int32_t Himg, int32_t Wimg, int32_t P, int32_t ws,
                       void* out16, int64_t ldo, int32_t Kpad, int32_t dtype, void* stream);
/* ViT output permute to NCHW (vitdet.py:847-857) and its gradient. */
int sam3b_tokens_to_nchw(const float* x, int32_t B, int32_t G, int32_t ws, int32_t D, float* out, void* stream);
int sam3b_nchw_to_tokens(const float* g, int32_t B, int32_t G, int32_t ws, int32_t D, float* dx, void* dx16,
                         int64_t ld16, int32_t dtype, void* stream);

/* ---- LoRA operand packing (lora_layers.py:39-47 parameter layout: A [in][r], B [r][out]) ----- */
typedef struct sam3b_lora_site {
  int32_t in, out_total, n, r, rpad;
  int32_t out_off[3], out_len[3];
  const float* A[3];
  const float* B[3];
} sam3b_lora_site;
int sam3b_lora_pack(const sam3b_lora_site* site, void* down_T, void* w_ext, int64_t ldw, void* up_pack, void* wt_ext,
                    int64_t ldwt, int32_t dtype, void* stream);
int sam3b_lora_unpack_grads(const sam3b_lora_site* site, const float* dA_pack, const float* dB_pack, float* const* dA,
                            float* const* dB, void* stream);

/* torch.optim.AdamW step on a flat fp32 buffer (train_sam3_lora_native.py:736-740); g is scaled by grad_scale first */
int sam3b_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAM3B_H_ */
