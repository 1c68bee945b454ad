// extern "C" surface of libsam3b.so (declared in include/sam3b.h).
#include "../../include/sam3b.h"

#include "common.h"
#include "attn.cuh"
#include "conv.cuh"
#include "elementwise.cuh"
#include "engine.h"
#include "gemm.cuh"
#include "input.cuh"
#include "loss.cuh"
#include "matcher.cuh"

#ifdef SAM3B_TRACE
namespace sam3b { int attn_trace_read(unsigned long long*, int); int attn_trace_clear(); int attn_trace_read_fwd(unsigned long long*, int); }
#endif

using namespace sam3b;

extern "C" {

const char* sam3b_last_error(void) { return last_error_message(); }
int sam3b_abi_version(void) { return SAM3B_ABI_VERSION; }
int64_t sam3b_launch_count(void) { return launch_count(); }

int sam3b_gemm(const sam3b_gemm_desc* d, void* stream) {
  if (!d) return fail(-1, "sam3b_gemm: null descriptor");
  GemmArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K;
  a.A = d->A; a.lda = d->lda; a.a_mn = d->a_mn;
  a.B = d->B; a.ldb = d->ldb; a.b_mn = d->b_mn;
  a.dtype = d->dtype; a.epilogue = d->epilogue;
  a.C = d->C; a.ldc = d->ldc; a.C2 = d->C2; a.ldc2 = d->ldc2;
  a.bias = d->bias;
  a.residual = d->residual; a.ldres = d->ldres; a.res_row_mod = d->res_row_mod;
  a.row_scale = d->row_scale; a.rows_per_scale = d->rows_per_scale;
  a.drop_p = d->drop_p; a.drop_seed = d->drop_seed;
  a.aux = d->aux; a.ldaux = d->ldaux;
  a.rope = d->rope; a.rope_period = d->rope_period; a.rope_cols = d->rope_cols;
  a.alpha = d->alpha; a.splitk = d->splitk; a.c_trans = d->c_trans; a.bn = d->bn;
  a.dbg_lbo = d->dbg_lbo; a.dbg_sbo = d->dbg_sbo; a.max_ctas = d->max_ctas; a.cta_pair = d->cta_pair;
  return gemm_launch(a, static_cast<cudaStream_t>(stream));
}

int sam3b_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t D,
                        void* y16, int64_t ldy, int32_t dtype, float* mean, float* rstd, void* stream) {
  return layernorm_fwd(x, gamma, beta, eps, rows, D, y16, ldy, dtype, mean, rstd, static_cast<cudaStream_t>(stream));
}
int sam3b_layernorm_bwd(const void* dy16, int64_t lddy, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* dres, int32_t rows, int32_t D, float* dx, void* dx16,
                        int64_t lddx16, int32_t dtype, void* stream) {
  return layernorm_bwd(dy16, lddy, x, mean, rstd, gamma, dres, rows, D, dx, dx16, lddx16, dtype,
                       static_cast<cudaStream_t>(stream));
}
int sam3b_cast_rows_16(const float* x, int32_t rows, int32_t D, void* y16, int64_t ldy, int32_t dtype, void* stream) {
  return cast_rows_16(x, rows, D, y16, ldy, dtype, static_cast<cudaStream_t>(stream));
}
int sam3b_cast_rows_16_scaled(const float* x, int32_t rows, int32_t D, void* y16, int64_t ldy, int32_t dtype, const float* scale,
                              void* stream) {
  return cast_rows_16(x, rows, D, y16, ldy, dtype, static_cast<cudaStream_t>(stream), scale);
}

static AttnArgs to_attn_args(const sam3b_attn_desc* d) {
  AttnArgs a;
  a.q = d->q; a.ldq = d->ldq; a.q_cols = d->q_cols; a.q_col0 = d->q_col0;
  a.kv = d->kv; a.ldkv = d->ldkv; a.kv_cols = d->kv_cols; a.k_col0 = d->k_col0; a.v_col0 = d->v_col0;
  a.nseg = d->nseg; a.Lq = d->Lq; a.Lk = d->Lk; a.heads = d->heads; a.dtype = d->dtype; a.scale = d->scale;
  a.O = d->O; a.ldo = d->ldo; a.o_col0 = d->o_col0; a.lse2 = d->lse2;
  a.bias = d->bias; a.kpm = d->kpm; a.drop_p = d->drop_p; a.drop_seed = d->drop_seed;
  a.dO = d->dO; a.lddo = d->lddo; a.do_col0 = d->do_col0; a.delta = d->delta;
  a.dq = d->dq; a.lddq = d->lddq; a.dq_col0 = d->dq_col0;
  a.dkv = d->dkv; a.lddkv = d->lddkv; a.dk_col0 = d->dk_col0; a.dv_col0 = d->dv_col0;
  a.rope = d->rope; a.rope_period = d->rope_period;
  a.drop_bits = d->drop_bits; a.drop_bitsT = d->drop_bitsT;
  return a;
}
int sam3b_attention_dropout_bits(int32_t n_bh, int32_t Lq, int32_t Lk, float p, uint32_t seed, uint32_t* bits, uint32_t* bitsT,
                                 void* stream) {
  return attn_dropout_bits(n_bh, Lq, Lk, p, seed, bits, bitsT, static_cast<cudaStream_t>(stream));
}
int sam3b_attention_fwd(const sam3b_attn_desc* d, void* stream) {
  if (!d) return fail(-1, "sam3b_attention_fwd: null descriptor");
  return attn_fwd_launch(to_attn_args(d), static_cast<cudaStream_t>(stream));
}
int sam3b_attention_bwd(const sam3b_attn_desc* d, void* stream) {
  if (!d) return fail(-1, "sam3b_attention_bwd: null descriptor");
  if (!d->O || !d->dO || !d->delta) return fail(-1, "sam3b_attention_bwd: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // delta = rowsum(dO * O) per (row, head), in the statistics layout of lse2
  const uint16_t* O16 = static_cast<const uint16_t*>(d->O) + d->o_col0;
  const uint16_t* dO16 = static_cast<const uint16_t*>(d->dO) + d->do_col0;
  int rc = attn_delta(dO16, d->lddo, O16, d->ldo, d->nseg * d->Lq, d->heads, d->dtype, d->delta, st, d->Lq);
  if (rc) return rc;
  return attn_bwd_launch(to_attn_args(d), st);
}

int sam3b_patch_gather(const float* img, int32_t B, int32_t C, int32_t Himg, int32_t Wimg, int32_t P, int32_t ws,
                       void* out16, int64_t ldo, int32_t Kpad, int32_t dtype, void* stream) {
  return patch_gather(img, B, C, Himg, Wimg, P, ws, out16, ldo, Kpad, dtype, static_cast<cudaStream_t>(stream));
}
int sam3b_tokens_to_nchw(const float* x, int32_t B, int32_t G, int32_t ws, int32_t D, float* out, void* stream) {
  return tokens_to_nchw(x, B, G, ws, D, out, static_cast<cudaStream_t>(stream));
}
int sam3b_nchw_to_tokens(const float* g, int32_t B, int32_t G, int32_t ws, int32_t D, float* dx, void* dx16,
                         int64_t ld16, int32_t dtype, void* stream) {
  return nchw_to_tokens(g, B, G, ws, D, dx, dx16, ld16, dtype, static_cast<cudaStream_t>(stream));
}

static LoraSite to_site(const sam3b_lora_site* s) {
  LoraSite o;
  o.in = s->in; o.out_total = s->out_total; o.n = s->n; o.r = s->r; o.rpad = s->rpad;
  for (int i = 0; i < 3; ++i) { o.out_off[i] = s->out_off[i]; o.out_len[i] = s->out_len[i]; o.A[i] = s->A[i]; o.B[i] = s->B[i]; }
  return o;
}
int sam3b_lora_pack(const sam3b_lora_site* site, void* down_T, void* w_ext, int64_t ldw, void* up_pack, void* wt_ext,
                    int64_t ldwt, int32_t dtype, void* stream) {
  if (!site) return fail(-1, "sam3b_lora_pack: null site");
  return lora_pack(to_site(site), down_T, w_ext, ldw, up_pack, wt_ext, ldwt, dtype, static_cast<cudaStream_t>(stream));
}
int sam3b_lora_unpack_grads(const sam3b_lora_site* site, const float* dA_pack, const float* dB_pack, float* const* dA,
                            float* const* dB, void* stream) {
  if (!site || !dA || !dB) return fail(-1, "sam3b_lora_unpack_grads: null argument");
  float* a3[3] = {nullptr, nullptr, nullptr};
  float* b3[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < site->n && i < 3; ++i) { a3[i] = dA[i]; b3[i] = dB[i]; }
  return lora_unpack_grads(to_site(site), dA_pack, dB_pack, a3, b3, static_cast<cudaStream_t>(stream));
}
int sam3b_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t step, float grad_scale, void* stream) {
  return adamw_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, static_cast<cudaStream_t>(stream));
}

struct sam3b_vit { VitEngine* eng; };

int sam3b_vit_create(const sam3b_vit_config* c, sam3b_vit** out) {
  if (!c || !out) return fail(-1, "sam3b_vit_create: null argument");
  if (c->n_global < 0 || c->n_global > 16) return fail(-1, "sam3b_vit_create: n_global %d", c->n_global);
  if (c->embed_dim != c->num_heads * 64) return fail(-1, "sam3b_vit_create: head_dim must be 64");
  if (c->img_size % c->patch_size != 0 || (c->img_size / c->patch_size) % c->window_size != 0)
    return fail(-1, "sam3b_vit_create: img/patch/window sizes do not tile");
  if (c->lora_rank < 0 || c->lora_rank > 64) return fail(-1, "sam3b_vit_create: lora_rank %d outside [0,64]", c->lora_rank);
  if ((c->window_size * c->window_size) % 64 != 0) return fail(-1, "sam3b_vit_create: window tokens must be a multiple of 64");
  if (c->mlp_hidden % 8 != 0) return fail(-1, "sam3b_vit_create: mlp_hidden %% 8 != 0");
  VitConfig v;
  v.img_size = c->img_size; v.patch_size = c->patch_size; v.in_chans = c->in_chans; v.embed_dim = c->embed_dim;
  v.depth = c->depth; v.num_heads = c->num_heads; v.mlp_hidden = c->mlp_hidden; v.window_size = c->window_size;
  v.global_blocks.assign(c->global_blocks, c->global_blocks + c->n_global);
  v.pos_side = c->pos_side; v.ln_eps = c->ln_eps; v.rope_theta = c->rope_theta;
  v.lora_rank = c->lora_rank; v.lora_scaling = c->lora_scaling; v.lora_targets = c->lora_targets;
  v.dtype = c->dtype; v.max_batch = c->max_batch;
  *out = new sam3b_vit{new VitEngine(v)};
  return 0;
}
void sam3b_vit_destroy(sam3b_vit* v) {
  if (v) { delete v->eng; delete v; }
}
int64_t sam3b_vit_weight_bytes(const sam3b_vit* v) { return v ? v->eng->weight_bytes() : -1; }
int64_t sam3b_vit_workspace_bytes(sam3b_vit* v, int32_t batch, int32_t training) {
  return v ? v->eng->workspace_bytes(batch, training != 0) : -1;
}
int64_t sam3b_vit_lora_numel(const sam3b_vit* v) { return v ? v->eng->lora_numel() : -1; }
int32_t sam3b_vit_lora_count(const sam3b_vit* v) { return v ? (int32_t)v->eng->lora_entries().size() : -1; }
int sam3b_vit_lora_entry(const sam3b_vit* v, int32_t index, sam3b_lora_entry* out) {
  if (!v || !out) return fail(-1, "sam3b_vit_lora_entry: null argument");
  const auto& es = v->eng->lora_entries();
  if (index < 0 || index >= (int32_t)es.size()) return fail(-1, "sam3b_vit_lora_entry: index %d out of range", index);
  const LoraEntry& e = es[index];
  out->block = e.block; out->target = e.target; out->in = e.in; out->out = e.out; out->rank = e.rank;
  out->a_off = e.a_off; out->b_off = e.b_off;
  return 0;
}
int sam3b_vit_bind(sam3b_vit* v, void* weight_buf, int64_t weight_bytes, void* work_buf, int64_t work_bytes,
                   int32_t batch, int32_t training) {
  if (!v) return fail(-1, "sam3b_vit_bind: null handle");
  return v->eng->bind(weight_buf, weight_bytes, work_buf, work_bytes, batch, training != 0);
}
int sam3b_vit_load_base(sam3b_vit* v, const float* const* tensors, int32_t n_tensors, void* stream) {
  if (!v || !tensors) return fail(-1, "sam3b_vit_load_base: null argument");
  return v->eng->load_base(tensors, n_tensors, static_cast<cudaStream_t>(stream));
}
int sam3b_vit_forward(sam3b_vit* v, const float* img, int32_t batch, const float* lora_flat, float* out_nchw,
                      int32_t save_for_backward, void* stream) {
  if (!v || !img) return fail(-1, "sam3b_vit_forward: null argument");
  return v->eng->forward(img, batch, lora_flat, out_nchw, save_for_backward != 0, static_cast<cudaStream_t>(stream));
}
int sam3b_vit_set_drop_path(sam3b_vit* v, const float* scales) {
  if (!v) return fail(-1, "sam3b_vit_set_drop_path: null handle");
  v->eng->set_drop_path(scales);
  return 0;
}
int sam3b_vit_set_lora_dropout(sam3b_vit* v, float p, uint32_t seed) {
  if (!v) return fail(-1, "sam3b_vit_set_lora_dropout: null handle");
  return v->eng->set_lora_dropout(p, seed);
}
int sam3b_vit_set_lora_dropout_dev(sam3b_vit* v, float p, uint32_t seed, const uint32_t* seed_dev) {
  if (!v) return fail(-1, "sam3b_vit_set_lora_dropout_dev: null handle");
  return v->eng->set_lora_dropout(p, seed, seed_dev);
}
int sam3b_dropout_rows16(const void* x16, int64_t ldx, int32_t rows, int32_t cols, void* out16, int64_t ldo, float p,
                         uint32_t seed, int32_t dtype, void* stream) {
  return dropout_rows16(x16, ldx, rows, cols, out16, ldo, p, seed, dtype, static_cast<cudaStream_t>(stream));
}
int sam3b_vit_backward(sam3b_vit* v, const float* gout_nchw, float* lora_grad_flat, void* stream) {
  if (!v) return fail(-1, "sam3b_vit_backward: null handle");
  return v->eng->backward(gout_nchw, lora_grad_flat, static_cast<cudaStream_t>(stream));
}
int sam3b_vit_backward_segment(sam3b_vit* v, const float* gout_nchw, float* lora_grad_flat, int32_t block_hi, int32_t block_lo,
                               void* stream) {
  if (!v) return fail(-1, "sam3b_vit_backward_segment: null handle");
  return v->eng->backward_segment(gout_nchw, lora_grad_flat, block_hi, block_lo, static_cast<cudaStream_t>(stream));
}
int sam3b_vit_lora_grad_range(sam3b_vit* v, int32_t block_hi, int32_t block_lo, int64_t* lo, int64_t* hi) {
  if (!v || !lo || !hi) return fail(-1, "sam3b_vit_lora_grad_range: null argument");
  v->eng->lora_grad_range(block_hi, block_lo, lo, hi);
  return 0;
}

int sam3b_focal_loss_fwd(const float* x, const float* y, int64_t n, float alpha, float gamma, float* loss, float* sum, void* stream) {
  return focal_loss_fwd(x, y, n, alpha, gamma, loss, sum, static_cast<cudaStream_t>(stream));
}
int sam3b_focal_loss_bwd(const float* x, const float* y, int64_t n, float alpha, float gamma, const float* g, float gscale,
                         float* dx, void* stream) {
  return focal_loss_bwd(x, y, n, alpha, gamma, g, gscale, dx, static_cast<cudaStream_t>(stream));
}

int sam3b_mask_loss_fwd(const float* src, int32_t N, int32_t h, int32_t w, const void* tgt, int32_t tgt_u8, int32_t H, int32_t W,
                        float alpha, float gamma, float num_boxes, float* partial, float* sums, float* out, void* stream) {
  return mask_loss_fwd(src, N, h, w, tgt, tgt_u8, H, W, alpha, gamma, num_boxes, partial, sums, out, static_cast<cudaStream_t>(stream));
}
int sam3b_mask_loss_bwd(const float* src, int32_t N, int32_t h, int32_t w, const void* tgt, int32_t tgt_u8, int32_t H, int32_t W,
                        float alpha, float gamma, float num_boxes, const float* sums, const float* g, float* dsrc, void* stream) {
  return mask_loss_bwd(src, N, h, w, tgt, tgt_u8, H, W, alpha, gamma, num_boxes, sums, g, dsrc, static_cast<cudaStream_t>(stream));
}

int sam3b_matcher(const sam3b_matcher_desc* d, float* cost, int32_t* query_of_col, int32_t* col_of_query, void* stream) {
  if (!d) return fail(-1, "sam3b_matcher: null descriptor");
  MatcherArgs a;
  a.B = d->B; a.Q = d->Q; a.Tmax = d->Tmax; a.repeats = d->repeats;
  a.logits = d->logits; a.pred_boxes = d->pred_boxes; a.tgt_boxes = d->tgt_boxes; a.num_boxes = d->num_boxes;
  a.out_valid = d->out_valid; a.tgt_valid = d->tgt_valid;
  a.w_class = d->w_class; a.w_bbox = d->w_bbox; a.w_giou = d->w_giou;
  a.focal = d->focal; a.stable = d->stable; a.alpha = d->alpha; a.gamma = d->gamma;
  return matcher_run(a, cost, query_of_col, col_of_query, static_cast<cudaStream_t>(stream));
}

int sam3b_resample_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* coeffs) {
  return resample_coeffs(in_size, out_size, bounds, coeffs);
}
int sam3b_image_resize_normalize(const uint8_t* src, int32_t h, int32_t w, int32_t out, const int32_t* bounds_x, const int32_t* coeffs_x,
                                 int32_t ks_x, const int32_t* bounds_y, const int32_t* coeffs_y, int32_t ks_y, uint8_t* tmp, float* dst,
                                 float mean, float std, void* stream) {
  return image_resize_normalize(src, h, w, out, bounds_x, coeffs_x, ks_x, bounds_y, coeffs_y, ks_y, tmp, dst, mean, std,
                                static_cast<cudaStream_t>(stream));
}
int sam3b_rle_masks_nearest(const uint32_t* cum, const int32_t* offs, const int32_t* hw, int32_t N, int32_t out, uint8_t* dst,
                            void* stream) {
  return rle_masks_nearest(cum, offs, hw, N, out, dst, static_cast<cudaStream_t>(stream));
}
int sam3b_poly_crossings(const int32_t* edges, const int32_t* pt_start, int32_t n_edges, int64_t total_pts, int64_t* keys, void* stream) {
  return poly_crossings(edges, pt_start, n_edges, total_pts, keys, static_cast<cudaStream_t>(stream));
}
int sam3b_poly_masks_nearest(const int64_t* keys_sorted, int64_t n_keys, const int32_t* list_ofs, const int32_t* hw, int32_t N,
                             int32_t out, uint8_t* dst, void* stream) {
  return poly_masks_nearest(keys_sorted, n_keys, list_ofs, hw, N, out, dst, static_cast<cudaStream_t>(stream));
}

#define SAM3B_ST static_cast<cudaStream_t>(stream)
int sam3b_grad_scale(const float* g, int64_t n, float target, float* scale, void* stream) { return grad_scale(g, n, target, scale, SAM3B_ST); }
int sam3b_scale_cast(const void* in, int32_t tin, void* out, int32_t tout, int64_t n, int32_t dtype, const float* scale,
                     int32_t accumulate, void* stream) {
  return scale_cast(in, tin, out, tout, n, dtype, scale, accumulate, SAM3B_ST);
}
int sam3b_transpose_cast(const void* in, int32_t tin, void* out, int32_t tout, int32_t batch, int32_t R, int32_t C, int32_t dtype,
                         const float* scale, void* stream) {
  return transpose_cast(in, tin, out, tout, batch, R, C, dtype, scale, SAM3B_ST);
}
int sam3b_im2col3x3(const void* x16, int32_t B, int32_t H, int32_t W, int32_t C, void* out16, int64_t ldo, void* stream) {
  return im2col3x3(x16, B, H, W, C, out16, ldo, SAM3B_ST);
}
int sam3b_conv3x3_supported(int32_t H, int32_t W, int32_t C, int32_t Cout) { return conv3x3_supported(H, W, C, Cout) ? 1 : 0; }
int sam3b_conv3x3(const void* x16, int32_t B, int32_t H, int32_t W, int32_t C, const void* w9, int32_t Cout, const float* bias,
                  void* out, int64_t ldc, int32_t out_f32, int32_t dtype, void* stream) {
  ConvArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C; a.Cout = Cout;
  a.x16 = x16; a.w9 = w9; a.bias = bias; a.out = out; a.ldc = ldc; a.out_f32 = out_f32; a.dtype = dtype;
  return conv3x3_launch(a, SAM3B_ST);
}
int sam3b_pixel_shuffle2(const void* in16, int32_t B, int32_t H, int32_t W, int32_t C, int32_t gelu, void* out16, int32_t dtype,
                         void* stream) {
  return pixel_shuffle2(in16, B, H, W, C, gelu, out16, dtype, SAM3B_ST);
}
int sam3b_pixel_unshuffle2(const void* dy16, const void* h16, int32_t B, int32_t H, int32_t W, int32_t C, void* out16, int32_t dtype,
                           void* stream) {
  return pixel_unshuffle2(dy16, h16, B, H, W, C, out16, dtype, SAM3B_ST);
}
int sam3b_maxpool2_fwd(const void* x16, int32_t B, int32_t H, int32_t W, int32_t C, void* y16, int32_t dtype, void* stream) {
  return maxpool2_fwd(x16, B, H, W, C, y16, dtype, SAM3B_ST);
}
int sam3b_maxpool2_bwd(const void* x16, const void* dy16, int32_t B, int32_t H, int32_t W, int32_t C, const float* scale, float* dx32,
                       int32_t dtype, void* stream) {
  return maxpool2_bwd(x16, dy16, B, H, W, C, scale, dx32, dtype, SAM3B_ST);
}
int sam3b_upsample_add(const void* prev16, int32_t h, int32_t w, const void* cur16, int32_t B, int32_t H, int32_t W, int32_t C,
                       void* out16, int32_t dtype, void* stream) {
  return upsample_add(prev16, h, w, cur16, B, H, W, C, out16, dtype, SAM3B_ST);
}
int sam3b_upsample_add_bwd(const void* dout16, int32_t B, int32_t H, int32_t W, int32_t C, int32_t h, int32_t w, void* dprev16,
                           int32_t dtype, void* stream) {
  return upsample_add_bwd(dout16, B, H, W, C, h, w, dprev16, dtype, SAM3B_ST);
}
int sam3b_groupnorm_stats(const float* x, int32_t B, int32_t HW, int32_t C, int32_t G, float eps, double* work, float* stat,
                          void* stream) {
  return groupnorm_stats(x, B, HW, C, G, eps, work, stat, SAM3B_ST);
}
int sam3b_groupnorm_relu_fwd(const float* x, const float* stat, const float* gamma, const float* beta, int32_t B, int32_t HW,
                             int32_t C, int32_t G, void* y, int32_t out_f32, int32_t dtype, void* stream) {
  return groupnorm_relu_fwd(x, stat, gamma, beta, B, HW, C, G, y, out_f32, dtype, SAM3B_ST);
}
int sam3b_groupnorm_relu_bwd(const void* dy16, const float* x, const float* stat, const float* gamma, const float* beta, int32_t B,
                             int32_t HW, int32_t C, int32_t G, double* work, void* dx16, int32_t dtype, void* stream) {
  return groupnorm_relu_bwd(dy16, x, stat, gamma, beta, B, HW, C, G, work, dx16, dtype, SAM3B_ST);
}
#undef SAM3B_ST

#ifdef SAM3B_TRACE
int sam3b_debug_trace_read(unsigned long long* host, int n) { return attn_trace_read(host, n); }
int sam3b_debug_trace_clear(void) { return attn_trace_clear(); }
int sam3b_debug_trace_read_fwd(unsigned long long* host, int n) { return attn_trace_read_fwd(host, n); }
#endif

}  // extern "C"
