// Input pipeline kernels (next-row f4): the reference prepares every sample on the host with PIL and pycocotools
// (train_sam3_lora_native.py:101-108, 146-167); these kernels do the same arithmetic on the GPU, bit-exactly:
//   image: PILImage.resize((R, R), BILINEAR) -> ToTensor -> Normalize(mean, std)       uint8 HWC -> fp32 CHW
//   masks: mask_utils.decode(RLE) -> F.interpolate(mode="nearest") -> > 0.5            run lengths -> uint8 [R][R]
// Byte / index work, HBM-bound; no tensor cores involved.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

// Pillow's precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for the BILINEAR filter over a whole axis, on the HOST.
// bounds: [out_size][2] = (first source index, tap count); coeffs: [out_size][ksize] fixed point (22 fractional bits).
// Returns ksize (taps per output sample); pass coeffs == nullptr to query it.
int resample_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* coeffs);

// src [h][w][3] uint8 -> tmp [h][out][3] uint8 (horizontal pass, 8-bit intermediate like Pillow) -> dst [3][out][out] fp32 =
// ((u8 / 255) - mean) / std.  bounds / coeffs tables live in device memory (uploaded by the caller).
int image_resize_normalize(const uint8_t* src, int h, int w, int out, const int32_t* bounds_x, const int32_t* coeffs_x, int ks_x,
                           const int32_t* bounds_y, const int32_t* coeffs_y, int ks_y, uint8_t* tmp, float* dst, float mean,
                           float std, cudaStream_t s);

// N run-length masks -> dst [N][out][out] uint8 (0/1).  cum: concatenated CUMULATIVE run lengths of all masks (column-major
// runs starting with zeros, pycocotools rleDecode); offs [N+1]: each mask's slice of cum; hw [N][2] = (height, width).
int rle_masks_nearest(const uint32_t* cum, const int32_t* offs, const int32_t* hw, int N, int out, uint8_t* dst, cudaStream_t s);

// Polygon masks = pycocotools frPyObjects + merge + decode (common/maskApi.c rleFrPoly), train_sam3_lora_native.py:152-156.
// poly_crossings: edges [n_edges][8] int32 = (xs, ys, xe, ye) of one edge in 5x up-sampled integer coordinates (the host's
// (int)(5 * coord + .5)), polygon list id, image height, image width, 1 if it is the first edge of its polygon; pt_start
// [n_edges + 1]: prefix sums of max(|dx|, |dy|) + 1 (points per edge).  keys [total_pts] receives (list id << 32 | x * h + y)
// for every fill toggle and INT64_MAX elsewhere; the caller sorts keys ascending.
int poly_crossings(const int32_t* edges, const int32_t* pt_start, int n_edges, int64_t total_pts, int64_t* keys, cudaStream_t s);
// N objects -> dst [N][out][out] uint8 (0/1): object n is the union of polygon lists list_ofs[n] .. list_ofs[n+1]-1; a source
// pixel is inside a polygon when an odd number of its sorted crossings lie at or before it (column-major), then
// F.interpolate(mode="nearest") to out x out.  hw [N][2] = (height, width).
int poly_masks_nearest(const int64_t* keys_sorted, int64_t n_keys, const int32_t* list_ofs, const int32_t* hw, int N, int out,
                       uint8_t* dst, cudaStream_t s);

}  // namespace sam3b
