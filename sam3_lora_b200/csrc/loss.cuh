#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {
// loss (elementwise, may be null) and/or sum (scalar, zeroed by the call, may be null)
int focal_loss_fwd(const float* x, const float* y, int64_t n, float alpha, float gamma, float* loss, float* sum, cudaStream_t s);
// dx = dL/dx * gscale * (g ? g[i] : 1)
int focal_loss_bwd(const float* x, const float* y, int64_t n, float alpha, float gamma, const float* g, float gscale, float* dx,
                   cudaStream_t s);
// Fused bilinear up-sample + sigmoid focal + dice over matched masks (loss_fns.py:105-123, 126-176, 689-707).
//   src [N][h][w] fp32 logits, tgt [N][H][W] uint8 (tgt_u8) or fp32 in {0,1}
//   partial: N * ceil(H/8) * 4 floats of scratch;  sums [N][4] = (focal sum, sum sig*t, sum sig, sum t) (kept for the backward)
//   out[0] = loss_mask, out[1] = loss_dice (both already divided by num_boxes, as the reference returns them)
int mask_loss_fwd(const float* src, int N, int h, int w, const void* tgt, int tgt_u8, int H, int W, float alpha, float gamma,
                  float num_boxes, float* partial, float* sums, float* out, cudaStream_t s);
// dsrc [N][h][w] = g[0] * d loss_mask / d src + g[1] * d loss_dice / d src   (g: two device floats)
int mask_loss_bwd(const float* src, int N, int h, int w, const void* tgt, int tgt_u8, int H, int W, float alpha, float gamma,
                  float num_boxes, const float* sums, const float* g, float* dsrc, cudaStream_t s);
}  // namespace sam3b
