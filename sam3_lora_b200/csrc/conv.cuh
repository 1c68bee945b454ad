// HBM-bound kernels of the neck / pixel decoder / mask head (row a8 of the hot-path table: sam3/model/necks.py:100-125,
// sam3/model/maskformer_segmentation.py:23-51,203-219).  The convolutions themselves run on the tcgen05 GEMM (gemm.cu):
//   ConvTranspose2d(k=2,s=2)  = GEMM [pixels, Cin] x [4*Cout, Cin]^T, then pixel_shuffle2 (columns are (di, dj, co))
//   Conv2d 1x1                = GEMM
//   Conv2d 3x3, padding 1     = im2col3x3 (16-bit, k = (ky, kx, c)) + GEMM
//   einsum bqc,bchw->bqhw     = one GEMM per image
// Activations are channels-last (NHWC) 16-bit; GroupNorm statistics and all module outputs are fp32.
// The backward chains run on gradients multiplied by a device-resident power-of-two scale (grad_scale) so that fp16
// operands neither overflow nor flush to zero; the last kernel of a chain multiplies by 1/scale.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

// scale[0] = s = 2^floor(log2(target / max|g|)) (1 if g is all zero / not finite), scale[1] = 1/s, scale[2] = scratch.
int grad_scale(const float* g, int64_t n, float target, float* scale, cudaStream_t s);

// Element types below: 0 = 16-bit (fp16 / bf16 per `dtype`), 1 = fp32.
// out[i] = (tout)(in[i] * *scale)   (scale may be null);  accumulate != 0: out[i] += ... (fp32 out only).  n % 4 == 0.
int scale_cast(const void* in, int tin, void* out, int tout, int64_t n, int dtype, const float* scale, int accumulate,
               cudaStream_t s);
// in [batch][R][C] -> out [batch][C][R]  (NCHW <-> NHWC with R = C_channels / HW; 2-D transposes of the einsum backward)
int transpose_cast(const void* in, int tin, void* out, int tout, int batch, int R, int C, int dtype, const float* scale,
                   cudaStream_t s);

// x16 [B][H][W][C] -> out16 [B*H*W][ldo], column (ky*3+kx)*C + c = x[b][y+ky-1][x+kx-1][c] (zero outside).  C % 8 == 0.
int im2col3x3(const void* x16, int B, int H, int W, int C, void* out16, int64_t ldo, cudaStream_t s);

// in16 [B*H*W][4*C] with columns (di, dj, c) -> out16 [B][2H][2W][C]; gelu != 0 applies exact-erf GELU on the way.
int pixel_shuffle2(const void* in16, int B, int H, int W, int C, int gelu, void* out16, int dtype, cudaStream_t s);
// inverse (gradient path): out16 [B*H*W][4*C] = dy16 [B][2H][2W][C] (* gelu'(h16) when h16 != null, same layout as out16)
int pixel_unshuffle2(const void* dy16, const void* h16, int B, int H, int W, int C, void* out16, int dtype, cudaStream_t s);

// nn.MaxPool2d(2, 2) on NHWC 16-bit; backward recomputes the arg-max (first maximum in (0,0),(0,1),(1,0),(1,1) order,
// as ATen does) and ACCUMULATES *scale * dy into the fp32 NHWC gradient dx32.
int maxpool2_fwd(const void* x16, int B, int H, int W, int C, void* y16, int dtype, cudaStream_t s);
int maxpool2_bwd(const void* x16, const void* dy16, int B, int H, int W, int C, const float* scale, float* dx32, int dtype,
                 cudaStream_t s);

// out = cur + nearest_upsample(prev) (F.interpolate(mode="nearest") for integer factors H/h, W/w) and its adjoint
int upsample_add(const void* prev16, int h, int w, const void* cur16, int B, int H, int W, int C, void* out16, int dtype,
                 cudaStream_t s);
int upsample_add_bwd(const void* dout16, int B, int H, int W, int C, int h, int w, void* dprev16, int dtype, cudaStream_t s);

// GroupNorm(G, C) + ReLU over NHWC fp32 x [B][HW][C] (maskformer_segmentation.py:217).
//   groupnorm_stats: stat [B][G] (mean, rstd) fp32; `work` = B*G*2 doubles of scratch
//   groupnorm_relu_fwd: y = relu((x - mean) * rstd * gamma + beta), 16-bit or fp32 (out_f32)
//   groupnorm_relu_bwd: dx16 = d/dx of the above for upstream dy16 (recomputes the ReLU mask); `work` as above
int groupnorm_stats(const float* x, int B, int HW, int C, int G, float eps, double* work, float* stat, cudaStream_t s);
int groupnorm_relu_fwd(const float* x, const float* stat, const float* gamma, const float* beta, int B, int HW, int C, int G,
                       void* y, int out_f32, int dtype, cudaStream_t s);
int groupnorm_relu_bwd(const void* dy16, const float* x, const float* stat, const float* gamma, const float* beta, int B, int HW,
                       int C, int G, double* work, void* dx16, int dtype, cudaStream_t s);

}  // namespace sam3b
