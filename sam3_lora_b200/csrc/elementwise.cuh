// HBM-bound helper kernels of the ViT trunk (host-side launch interface).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

// y16[row][0..D) = LayerNorm(x[row]) * gamma + beta   (nn.LayerNorm eps, vitdet.py:566,584,719)
// Saves mean / rstd per row for the backward.  x: fp32 [rows][D] contiguous.
int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int rows, int D,
                  void* y16, int64_t ldy, int dtype, float* mean, float* rstd, cudaStream_t s);

// dx = dres + dLN(dy16; x, mean, rstd, gamma); also a 16-bit copy of dx for the next GEMM.
// dres may alias dx (in-place accumulate).  dx16 may be null.
// dx16 = row_scale[row / rows_per_scale] * dx when row_scale != null (DropPath scaling of the branch gradient).
int layernorm_bwd(const void* dy16, int64_t lddy, const float* x, const float* mean, const float* rstd,
                  const float* gamma, const float* dres, int rows, int D, float* dx, void* dx16, int64_t lddx16,
                  int dtype, cudaStream_t s, const float* row_scale = nullptr, int rows_per_scale = 1);

// y32[row][0..D) = LayerNorm(x[row]) * gamma + beta in fp32 (ln_pre feeding the fp32 residual stream,
// vitdet.py:833).  In-place (y32 == x) is allowed.
int layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float eps, int rows, int D, float* y32,
                      cudaStream_t s);

// Frozen-weight packing: W fp32 [N][K] -> 16-bit dst[n][k] (ld) or, transposed, dst[k][n] (ld).
int pack_weight(const float* W, int N, int K, void* dst16, int64_t ld, int transpose, int dtype, cudaStream_t s);

// Tiled absolute position table in window-major token order (get_abs_pos tiling, vitdet.py:199-222):
// out[t][:] = pos_embed[1 + (pi % side)*side + (pj % side)][:] for token t at patch (pi, pj).
int build_pos_table(const float* pos_embed, int side, int G, int ws, int D, float* out, cudaStream_t s);

// out16[row][c] = keep(row,c) ? x16[row][c] / (1-p) : 0  for c < cols (inverted dropout, mask from rng.cuh)
int dropout_rows16(const void* x16, int64_t ldx, int rows, int cols, void* out16, int64_t ldo, float p, uint32_t seed,
                   int dtype, cudaStream_t s, const uint32_t* seed_dev = nullptr);

// y16[row][0..D) = (16-bit) x[row][0..D)
// scale (device scalar, optional): y16 = (16-bit)(x * *scale)
int cast_rows_16(const float* x, int rows, int D, void* y16, int64_t ldy, int dtype, cudaStream_t s, const float* scale = nullptr);

// delta[h][row] = sum_c dO[row][h*64+c] * O[row][h*64+c]   (head-major)
// delta[h][seg*Lq_stat + i] for row = seg*Lq + i (Lq_stat = Lq rounded up to 64; pass Lq = rows for one flat segment)
int attn_delta(const void* dO, int64_t lddo, const void* O, int64_t ldo, int rows, int heads, int dtype, float* delta,
               cudaStream_t s, int Lq = 0);

// Patch gather for the k=s=P patch-embed conv (vitdet.py:323-336): image fp32 NCHW ->
// 16-bit rows [token][Kpad] with k = (c*P + u)*P + v, tokens in window-major order:
// token = ((b*nwin + wy*nwx + wx)*ws + i)*ws + j  for patch (wy*ws+i, wx*ws+j).
int patch_gather(const float* img, int B, int C, int Himg, int Wimg, int P, int ws, void* out16, int64_t ldo, int Kpad,
                 int dtype, cudaStream_t s);

// Window-major token stream [B][T][D] fp32 <-> NCHW feature map [B][D][G][G] (vitdet.py:847-857).
int tokens_to_nchw(const float* x, int B, int G, int ws, int D, float* out, cudaStream_t s);
// and back (gradient path): also emits the 16-bit copy used as the first dgrad operand.
// gscale (device scalar, optional) multiplies both outputs (gradient scaling of the trunk backward).
int nchw_to_tokens(const float* g, int B, int G, int ws, int D, float* dx, void* dx16, int64_t ld16, int dtype,
                   cudaStream_t s, const float* img_scale = nullptr, const float* gscale = nullptr);

// ---- LoRA packing -------------------------------------------------------------------------
// One adapted Linear (in -> out_total) with n adapters of rank r that each own the output
// slice [out_off[a], out_off[a]+out_len[a]) (q/k/v slices of the fused qkv; one slice otherwise).
struct LoraSite {
  int in = 0, out_total = 0, n = 0, r = 0, rpad = 64;
  int out_off[3] = {0, 0, 0};
  int out_len[3] = {0, 0, 0};
  const float* A[3] = {nullptr, nullptr, nullptr};  // [in][r]   (lora_layers.py:40)
  const float* B[3] = {nullptr, nullptr, nullptr};  // [r][out_len] (lora_layers.py:41)
};
// Fills the four 16-bit operand blocks the fused GEMMs read:
//   down_T  [rpad][in]            row a*r+j = A_a[:, j]                 (B operand of T = x.A)
//   w_ext   [out_total][ldw]      cols in + a*r+j = B_a[j][n-off_a]     (K-extension of W)
//   up_pack [rpad][out_total]     row a*r+j, col n = B_a[j][n-off_a]    (B operand of dT = dy.B^T)
//   wt_ext  [in][ldwt]            cols out_total + a*r+j = A_a[k][j]    (K-extension of W^T)
int lora_pack(const LoraSite& site, void* down_T, void* w_ext, int64_t ldw, void* up_pack, void* wt_ext,
              int64_t ldwt, int dtype, cudaStream_t s);
// dA_a[k][j] = dA_pack[k][a*r+j] ; dB_a[j][n] = dB_pack[a*r+j][off_a+n]   (fp32)
// out_scale (device scalar, optional) multiplies every element written.
int lora_unpack_grads(const LoraSite& site, const float* dA_pack, const float* dB_pack, float* const dA[3],
                      float* const dB[3], cudaStream_t s, const float* out_scale = nullptr);

// Batched forms for the trunk engine: one launch for ALL adapted Linears of the trunk instead of one per site (128 pack +
// 128 unpack launches and 256 memsets per step at depth 32).  Descriptors live in device memory (built once per bind);
// adapter parameters / gradients are addressed as offsets into the flat LoRA buffers.
struct LoraSiteDesc {
  int in, out_total, n, r, rpad;
  int out_off[3], out_len[3];
  int64_t a_off[3], b_off[3];      // element offsets of A_a [in][r] / B_a [r][out_len] in the flat buffer
  void *down_T, *w_ext, *up_pack, *wt_ext;
  int64_t ldw, ldwt;
  float *dA_pack, *dB_pack;        // this site's split-K accumulators ([in][rpad], [rpad][out_total]); null in inference layouts
};
int lora_pack_all(const LoraSiteDesc* dev_descs, int n_sites, int max_work, const float* flat, int dtype, cudaStream_t s);
int lora_unpack_all(const LoraSiteDesc* dev_descs, int n_sites, int max_work, float* grad_flat, const float* out_scale,
                    cudaStream_t s);

// Fused AdamW over a flat fp32 buffer (torch.optim.AdamW semantics, train_sam3_lora_native.py:736-740).
int adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, cudaStream_t s);

}  // namespace sam3b
