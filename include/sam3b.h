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

#define SAM3B_ABI_VERSION 3

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
#define SAM3B_EPI_STORE32 6      /* C32 = [row_scale[row/rows_per_scale] *] (alpha*acc (+bias)) */
#define SAM3B_EPI_ADDMASK16 7    /* C16 += dropout_mask/(1-p) * alpha*acc [* gelu_erf'(aux16)] */

const char* sam3b_last_error(void);
int sam3b_abi_version(void);
/* number of CUDA kernels this library has launched in this process (all of them are ours) */
int64_t sam3b_launch_count(void);

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
  int32_t cta_pair;                         /* 0 = default, 1 = single-CTA tiles, 2 = CTA-pair (cta_group::2) tiles */
  const float* row_scale; int32_t rows_per_scale; /* RESIDUAL_F32: out = res + row_scale[row/rows_per_scale]*(acc+bias) */
  float drop_p; uint32_t drop_seed;         /* ADDMASK16 */
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
/* same with a device-resident factor: y16 = (16-bit)(x * *scale)  (gradient scaling, see sam3b_grad_scale) */
int sam3b_cast_rows_16_scaled(const float* x, int32_t rows, int32_t D, void* y16, int64_t ldy, int32_t dtype, const float* scale,
                              void* stream);

/* ---- attention (head_dim 64; tokens in window-major order so a window/image is a row run) --- */
/* apply_rotary_enc + F.scaled_dot_product_attention, vitdet.py:68-90,485,502 (RoPE itself is the
 * SAM3B_EPI_QKV_ROPE epilogue of the qkv GEMM). */
typedef struct sam3b_attn_desc {
  /* 16-bit operands; head h of q at column q_col0 + 64*h of `q`, of k / v at k_col0 / v_col0 + 64*h of `kv`.
   * head_dim 64, or 32 stored zero-padded to 64 columns per head (nn.MultiheadAttention sites, E=256, 8 heads). */
  const void* q; int64_t ldq; int32_t q_cols, q_col0;
  const void* kv; int64_t ldkv; int32_t kv_cols, k_col0, v_col0;
  int32_t nseg, Lq, Lk, heads, dtype; /* nseg independent problems: q rows seg*Lq.., kv rows seg*Lk.. */
  float scale;                         /* logical head_dim^-0.5 */
  void* O; int64_t ldo; int32_t o_col0; /* fwd out / bwd in */
  float* lse2;                         /* [heads][nseg*Lq_stat], Lq_stat = Lq rounded up to 64; log2-domain LSE */
  /* optional: additive float attn_mask [nseg*heads][Lq][Lk], key_padding_mask [nseg][Lk] (non-zero = ignore),
   * dropout on the attention probabilities (stateless hash mask, csrc/rng.cuh) */
  const float* bias; const uint8_t* kpm; float drop_p; uint32_t drop_seed;
  /* backward only */
  const void* dO; int64_t lddo; int32_t do_col0;
  float* delta;                        /* [heads][nseg*Lq_stat] scratch: rowsum(dO*O), written by the call */
  void* dq; int64_t lddq; int32_t dq_col0;
  void* dkv; int64_t lddkv; int32_t dk_col0, dv_col0;
  const float* rope; int32_t rope_period; /* optional inverse RoPE on dq, dk (ViT); NULL otherwise */
  /* ABI v3, optional: keep-bits of the dropout mask precomputed by sam3b_attention_dropout_bits (same hash, so the same mask
   * as the inline evaluation); the three kernels then read one word per 32 scores instead of hashing each score three times */
  const uint32_t* drop_bits; const uint32_t* drop_bitsT;
} sam3b_attn_desc;
/* bits[(bh*Lq + q)*pitch(Lk) + k/32] bit (k & 31) and bitsT[(bh*Lk + k)*pitch(Lq) + q/32] bit (q & 31) = 1 where probability
 * (q, k) of problem bh = seg*heads + head is kept; pitch(L) = L/32 rounded up to a multiple of 8 words (32-byte rows), so
 * bits holds n_bh*Lq*pitch(Lk) words and bitsT n_bh*Lk*pitch(Lq); both 32-byte aligned.  Lq, Lk multiples of 32. */
int sam3b_attention_dropout_bits(int32_t n_bh, int32_t Lq, int32_t Lk, float p, uint32_t seed, uint32_t* bits, uint32_t* bitsT,
                                 void* stream);
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

/* ---- ViT trunk engine (ViT.forward / Block.forward / Attention.forward + autograd backward) --- */
/* sam3/model/vitdet.py:813-859, 597-613, 466-515; only the LoRA adapters receive gradients
 * (apply_lora_to_model freezes the rest, lora_layers.py:171-172). */
#define SAM3B_LORA_Q 1
#define SAM3B_LORA_K 2
#define SAM3B_LORA_V 4
#define SAM3B_LORA_O 8
#define SAM3B_LORA_FC1 16
#define SAM3B_LORA_FC2 32

typedef struct sam3b_vit_config {
  int32_t img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_hidden, window_size;
  int32_t n_global; int32_t global_blocks[16];
  int32_t pos_side;              /* pos_embed is [1][1 + pos_side^2][D] (cls slot first), tiled over the grid */
  float ln_eps, rope_theta;
  int32_t lora_rank; float lora_scaling; int32_t lora_targets; /* SAM3B_LORA_* bitmask */
  int32_t dtype;                 /* SAM3B_F16 | SAM3B_BF16 tensor-core operands; residual stream is fp32 */
  int32_t max_batch;
} sam3b_vit_config;

typedef struct sam3b_lora_entry {
  int32_t block, target, in, out, rank;
  int64_t a_off, b_off;          /* element offsets into the flat fp32 LoRA buffers: A [in][rank], B [rank][out] */
} sam3b_lora_entry;

typedef struct sam3b_vit sam3b_vit; /* opaque host object; owns no device memory */

int sam3b_vit_create(const sam3b_vit_config* cfg, sam3b_vit** out);
void sam3b_vit_destroy(sam3b_vit* v);
int64_t sam3b_vit_weight_bytes(const sam3b_vit* v);
int64_t sam3b_vit_workspace_bytes(sam3b_vit* v, int32_t batch, int32_t training);
int64_t sam3b_vit_lora_numel(const sam3b_vit* v);
int32_t sam3b_vit_lora_count(const sam3b_vit* v);
int sam3b_vit_lora_entry(const sam3b_vit* v, int32_t index, sam3b_lora_entry* out);
/* caller-owned, 1024-byte aligned device buffers */
int sam3b_vit_bind(sam3b_vit* v, void* weight_buf, int64_t weight_bytes, void* work_buf, int64_t work_bytes,
                   int32_t batch, int32_t training);
/* fp32 device tensors in reference state-dict order: patch_embed.proj.weight, pos_embed, ln_pre.weight,
 * ln_pre.bias, then per block norm1.{weight,bias}, attn.qkv.{weight,bias}, attn.proj.{weight,bias},
 * norm2.{weight,bias}, mlp.fc1.{weight,bias}, mlp.fc2.{weight,bias}  (4 + 12*depth pointers, host array) */
int sam3b_vit_load_base(sam3b_vit* v, const float* const* tensors, int32_t n_tensors, void* stream);
/* img: fp32 NCHW [batch][C][S][S]; lora_flat: flat fp32 adapters; out: fp32 NCHW [batch][D][G][G] */
int sam3b_vit_forward(sam3b_vit* v, const float* img, int32_t batch, const float* lora_flat, float* out_nchw,
                      int32_t save_for_backward, void* stream);
/* DropPath (vitdet.py:610-611, rates linspace(0, 0.1, depth), model_builder.py:80): device array [depth][2][batch]
 * of per-image branch scales (0 or 1/keep; [i][0] attention, [i][1] MLP) for the next forward + backward; NULL = off */
int sam3b_vit_set_drop_path(sam3b_vit* v, const float* scales);
/* Adapter dropout (nn.Dropout on the LoRA branch input only, lora_layers.py:43,54) for the next forward + backward.
 * The mask is a stateless hash of (seed, block, site, row, col) — csrc/rng.cuh — regenerated where needed. p = 0: off */
int sam3b_vit_set_lora_dropout(sam3b_vit* v, float p, uint32_t seed);
/* same, with a device-resident word that the kernels ADD to `seed`: a captured CUDA graph then draws a fresh mask on every
 * replay if the caller rewrites *seed_dev between replays (the forward and its backward must see the same value) */
int sam3b_vit_set_lora_dropout_dev(sam3b_vit* v, float p, uint32_t seed, const uint32_t* seed_dev);
/* out16 = inverted-dropout(x16) with that mask (exposed for tests of the mask definition) */
int sam3b_dropout_rows16(const void* x16, int64_t ldx, int32_t rows, int32_t cols, void* out16, int64_t ldo, float p,
                         uint32_t seed, int32_t dtype, void* stream);
/* gout: dLoss/dout fp32 NCHW; lora_grad_flat: flat fp32 gradients (overwritten, same layout as lora_flat) */
int sam3b_vit_backward(sam3b_vit* v, const float* gout_nchw, float* lora_grad_flat, void* stream);
/* The same backward cut into consecutive block ranges [block_hi .. block_lo] (descending; the first call starts at
 * depth - 1 and is the only one that reads gout).  When a call returns (stream order), the gradients of ITS blocks are
 * final in elements [lo, hi) of lora_grad_flat (sam3b_vit_lora_grad_range), so a data-parallel caller can all-reduce that
 * slice while the next range runs — DDP's bucketed overlap (sam3_lora/train/native_trainer.py:322-340) without buckets. */
int sam3b_vit_backward_segment(sam3b_vit* v, const float* gout_nchw, float* lora_grad_flat, int32_t block_hi, int32_t block_lo,
                               void* stream);
int sam3b_vit_lora_grad_range(sam3b_vit* v, int32_t block_hi, int32_t block_lo, int64_t* lo, int64_t* hi);

/* ---- fused sigmoid focal loss (replaces the Triton kernels sam3/train/loss/sigmoid_focal_loss.py:35-208; arithmetic of
 * sam3/train/loss/loss_fns.py:159-167).  fp32, n elements; `loss` (elementwise) and `sum` (scalar) are optional outputs. */
int sam3b_focal_loss_fwd(const float* x, const float* y, int64_t n, float alpha, float gamma, float* loss, float* sum, void* stream);
/* dx[i] = dL/dx[i] * gscale * (g ? g[i] : 1) */
int sam3b_focal_loss_bwd(const float* x, const float* y, int64_t n, float alpha, float gamma, const float* g, float gscale,
                         float* dx, void* stream);

/* ---- fused mask losses: bilinear up-sampling (align_corners=False) of the matched mask logits to the target size + sigmoid
 * focal + dice in one pass over the targets (sam3/train/loss/loss_fns.py:105-123, 126-176, 689-707).  src [N][h][w] fp32, tgt
 * [N][H][W] uint8 (tgt_u8 != 0) or fp32; partial: N*ceil(H/8)*4 floats scratch; sums [N][4] kept for the backward;
 * out[0] = loss_mask, out[1] = loss_dice (already / num_boxes). */
int sam3b_mask_loss_fwd(const float* src, int32_t N, int32_t h, int32_t w, const void* tgt, int32_t tgt_u8, int32_t H, int32_t W,
                        float alpha, float gamma, float num_boxes, float* partial, float* sums, float* out, void* stream);
/* dsrc = g[0] * d loss_mask/d src + g[1] * d loss_dice/d src; g: two device floats (no host read-back) */
int sam3b_mask_loss_bwd(const float* src, int32_t N, int32_t h, int32_t w, const void* tgt, int32_t tgt_u8, int32_t H, int32_t W,
                        float alpha, float gamma, float num_boxes, const float* sums, const float* g, float* dsrc, void* stream);

/* ---- GPU-resident Hungarian matcher: BinaryHungarianMatcherV2.forward + _do_matching (sam3/train/matcher.py:15-29, 431-668)
 * without the .cpu().numpy() copy and the per-image scipy.optimize.linear_sum_assignment call. */
typedef struct sam3b_matcher_desc {
  int32_t B, Q, Tmax, repeats;
  const float* logits;          /* [B][Q] */
  const float* pred_boxes;      /* [B][Q][4] cxcywh */
  const float* tgt_boxes;       /* [B][Tmax][4] cxcywh (boxes_padded) */
  const int32_t* num_boxes;     /* [B] */
  const uint8_t* out_valid;     /* [B][Q] or NULL */
  const uint8_t* tgt_valid;     /* [B][Tmax] or NULL */
  float w_class, w_bbox, w_giou;
  int32_t focal, stable;
  float alpha, gamma;
} sam3b_matcher_desc;
/* cost [B][Q][Tmax] fp32 (output);  query_of_col [B][max(1,Tmax*repeats)] and col_of_query [B][Q] (int32, -1 = unmatched);
 * column c of image b is target c % num_boxes[b] (np.tile of the cost matrix when repeats > 1) */
int sam3b_matcher(const sam3b_matcher_desc* d, float* cost, int32_t* query_of_col, int32_t* col_of_query, void* stream);

/* ---- input pipeline (next-row f4): the reference's host-side sample preparation (train_sam3_lora_native.py:101-108, 146-167)
 * on the GPU, bit-exactly: PILImage.resize(BILINEAR) + ToTensor + Normalize, and RLE decode + nearest resize + > 0.5. */
/* HOST function: Pillow's coefficient tables for one axis (bounds [out][2], coeffs [out][ksize] int32); returns ksize
 * (call with NULL tables to query it), negative on error. */
int sam3b_resample_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* coeffs);
/* src [h][w][3] uint8 (device) -> dst [3][out][out] fp32; tmp: h*out*3 bytes of scratch; tables in device memory */
int sam3b_image_resize_normalize(const uint8_t* src, int32_t h, int32_t w, int32_t out, const int32_t* bounds_x, const int32_t* coeffs_x,
                                 int32_t ks_x, const int32_t* bounds_y, const int32_t* coeffs_y, int32_t ks_y, uint8_t* tmp, float* dst,
                                 float mean, float std, void* stream);
/* N RLE masks (cumulative run lengths, column-major, first run = zeros) -> dst [N][out][out] uint8 in {0, 1} */
int sam3b_rle_masks_nearest(const uint32_t* cum, const int32_t* offs, const int32_t* hw, int32_t N, int32_t out, uint8_t* dst,
                            void* stream);
/* Polygon masks: pycocotools frPyObjects + merge + decode (common/maskApi.c rleFrPoly) + nearest resize, the polygon branch
 * of train_sam3_lora_native.py:152-163.  Step 1: every point of the 5x up-sampled boundary walk decides whether it toggles
 * the column-major fill; edges [n_edges][8] int32 = (xs, ys, xe, ye, polygon list id, image h, image w, first-edge flag),
 * pt_start [n_edges + 1] prefix sums of points per edge; keys [total_pts] = (list id << 32 | x*h + y) or INT64_MAX.
 * The caller sorts keys.  Step 2: object n = union of lists list_ofs[n] .. list_ofs[n+1]-1, parity fill, nearest resize. */
int sam3b_poly_crossings(const int32_t* edges, const int32_t* pt_start, int32_t n_edges, int64_t total_pts, int64_t* keys, void* stream);
int sam3b_poly_masks_nearest(const int64_t* keys_sorted, int64_t n_keys, const int32_t* list_ofs, const int32_t* hw, int32_t N,
                             int32_t out, uint8_t* dst, void* stream);

/* ---- neck / pixel decoder / mask head helpers (row a8: sam3/model/necks.py:100-125, maskformer_segmentation.py:23-51,
 * 203-219).  The convolutions run on sam3b_gemm; these are the HBM-bound kernels around it.  Activations: channels-last
 * (NHWC) 16-bit; element-type codes below: 0 = 16-bit (per `dtype`), 1 = fp32. */
/* scale[0] = 2^floor(log2(target/max|g|)) (1 if g == 0), scale[1] = 1/scale[0], scale[2] = scratch (device floats) */
int sam3b_grad_scale(const float* g, int64_t n, float target, float* scale, void* stream);
/* out[i] = (tout)(in[i] * *scale) (scale may be NULL); accumulate: out[i] += (fp32 out only); n % 4 == 0 */
int sam3b_scale_cast(const void* in, int32_t tin, void* out, int32_t tout, int64_t n, int32_t dtype, const float* scale,
                     int32_t accumulate, void* stream);
/* in [batch][R][C] -> out [batch][C][R] (NCHW <-> NHWC; replaces the .permute()/.contiguous() pairs around the reference's convs) */
int sam3b_transpose_cast(const void* in, int32_t tin, void* out, int32_t tout, int32_t batch, int32_t R, int32_t C, int32_t dtype,
                         const float* scale, void* stream);
/* 3x3 / padding 1 patch matrix: out16[(b,y,x)][(ky*3+kx)*C + c] = x16[b][y+ky-1][x+kx-1][c] (nn.Conv2d(.,.,3,padding=1), necks.py:84-92) */
int sam3b_im2col3x3(const void* x16, int32_t B, int32_t H, int32_t W, int32_t C, void* out16, int64_t ldo, void* stream);
/* Implicit-GEMM nn.Conv2d(C, Cout, 3, padding=1) on channels-last 16-bit x16 [B][H][W][C] (necks.py:84-92,
 * maskformer_segmentation.py:187): no patch matrix, each filter tap's operand tile is one zero-filled 4-D TMA box.
 * w9 [Cout][9*C] with k = (ky*3+kx)*C + c; out [B*H*W][ldc] fp32 (out_f32) or 16-bit; bias [Cout] or NULL.
 * Needs C % 64 == 0, W % 8 == 0, Cout % 8 == 0 (sam3b_conv3x3_supported); otherwise use sam3b_im2col3x3 + sam3b_gemm. */
int sam3b_conv3x3_supported(int32_t H, int32_t W, int32_t C, int32_t Cout);
int sam3b_conv3x3(const void* x16, int32_t B, int32_t H, int32_t W, int32_t C, const void* w9, int32_t Cout, const float* bias,
                  void* out, int64_t ldc, int32_t out_f32, int32_t dtype, void* stream);
/* nn.ConvTranspose2d(k=2,s=2) output placement (necks.py:44-62): in16 [B*H*W][4C] columns (di,dj,c) -> out16 [B][2H][2W][C], optional GELU */
int sam3b_pixel_shuffle2(const void* in16, int32_t B, int32_t H, int32_t W, int32_t C, int32_t gelu, void* out16, int32_t dtype,
                         void* stream);
int sam3b_pixel_unshuffle2(const void* dy16, const void* h16, int32_t B, int32_t H, int32_t W, int32_t C, void* out16, int32_t dtype,
                           void* stream);
/* nn.MaxPool2d(2,2) (necks.py:66-69); the backward accumulates *scale * dy at the arg-max into the fp32 NHWC gradient */
int sam3b_maxpool2_fwd(const void* x16, int32_t B, int32_t H, int32_t W, int32_t C, void* y16, int32_t dtype, void* stream);
int sam3b_maxpool2_bwd(const void* x16, const void* dy16, int32_t B, int32_t H, int32_t W, int32_t C, const float* scale, float* dx32,
                       int32_t dtype, void* stream);
/* curr + F.interpolate(prev, size=curr.shape[-2:], mode="nearest") (maskformer_segmentation.py:208-210), integer factors */
int sam3b_upsample_add(const void* prev16, int32_t h, int32_t w, const void* cur16, int32_t B, int32_t H, int32_t W, int32_t C,
                       void* out16, int32_t dtype, void* stream);
int sam3b_upsample_add_bwd(const void* dout16, int32_t B, int32_t H, int32_t W, int32_t C, int32_t h, int32_t w, void* dprev16,
                           int32_t dtype, void* stream);
/* F.relu(GroupNorm(G, C)(x)) on NHWC fp32 x [B][HW][C] (maskformer_segmentation.py:216-217).  work: 3*B*G doubles of scratch;
 * stat: [B][G][2] fp32 (mean, rstd) */
int sam3b_groupnorm_stats(const float* x, int32_t B, int32_t HW, int32_t C, int32_t G, float eps, double* work, float* stat,
                          void* stream);
int sam3b_groupnorm_relu_fwd(const float* x, const float* stat, const float* gamma, const float* beta, int32_t B, int32_t HW,
                             int32_t C, int32_t G, void* y, int32_t out_f32, int32_t dtype, void* stream);
int sam3b_groupnorm_relu_bwd(const void* dy16, const float* x, const float* stat, const float* gamma, const float* beta, int32_t B,
                             int32_t HW, int32_t C, int32_t G, double* work, void* dx16, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAM3B_H_ */
