// extern "C" surface of libsam3b.so (declared in include/sam3b.h).
#include "../../include/sam3b.h"

#include "common.h"
#include "attn.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"

using namespace sam3b;

extern "C" {

const char* sam3b_last_error(void) { return last_error_message(); }
int sam3b_abi_version(void) { return SAM3B_ABI_VERSION; }

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
  a.aux = d->aux; a.ldaux = d->ldaux;
  a.rope = d->rope; a.rope_period = d->rope_period; a.rope_cols = d->rope_cols;
  a.alpha = d->alpha; a.splitk = d->splitk; a.c_trans = d->c_trans; a.bn = d->bn;
  a.dbg_lbo = d->dbg_lbo; a.dbg_sbo = d->dbg_sbo; a.max_ctas = d->max_ctas;
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

int sam3b_attention_fwd(const sam3b_attn_desc* d, void* stream) {
  if (!d) return fail(-1, "sam3b_attention_fwd: null descriptor");
  AttnFwdArgs a;
  a.qkv = d->qkv; a.ldqkv = d->ldqkv; a.tokens = d->tokens; a.seg_len = d->seg_len; a.D = d->D; a.heads = d->heads;
  a.head_dim = d->head_dim; a.dtype = d->dtype; a.O = d->O; a.ldo = d->ldo; a.lse2 = d->lse2;
  if (!a.qkv || !a.O || !a.lse2) return fail(-1, "sam3b_attention_fwd: null tensor");
  return attn_fwd_launch(a, static_cast<cudaStream_t>(stream));
}
int sam3b_attention_bwd(const sam3b_attn_desc* d, void* stream) {
  if (!d) return fail(-1, "sam3b_attention_bwd: null descriptor");
  if (!d->qkv || !d->O || !d->lse2 || !d->dO || !d->delta || !d->dqkv) return fail(-1, "sam3b_attention_bwd: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = attn_delta(d->dO, d->lddo, d->O, d->ldo, d->tokens, d->heads, d->dtype, d->delta, st);
  if (rc) return rc;
  AttnBwdArgs a;
  a.qkv = d->qkv; a.ldqkv = d->ldqkv; a.dO = d->dO; a.lddo = d->lddo; a.lse2 = d->lse2; a.delta = d->delta;
  a.dqkv = d->dqkv; a.lddqkv = d->lddqkv; a.rope = d->rope; a.rope_period = d->rope_period;
  a.tokens = d->tokens; a.seg_len = d->seg_len; a.D = d->D; a.heads = d->heads; a.head_dim = d->head_dim; a.dtype = d->dtype;
  return attn_bwd_launch(a, st);
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

}  // extern "C"
