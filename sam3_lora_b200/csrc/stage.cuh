// Warp-private shared-memory staging tile for coalesced row I/O.  tcgen05.ld 32x32b hands every lane one ROW, so a direct
// 16-byte access per lane touches 32 different 128-byte lines per instruction (32 LSU wavefronts).  A warp therefore owns
// a 2 KB tile (32 rows x 64 bytes): lanes write / read their own row segment, and global memory is accessed with 4 lanes
// per row, so one instruction covers 8 rows x 64 contiguous bytes.  The 16-byte chunks are XOR-swizzled by (row >> 1) & 3,
// which is conflict-free for both access patterns.  Used by the GEMM epilogues (gemm.cu) and by the attention backward's
// item transitions (attn_bwd.cu: parking the stationary rows, inverse-RoPE table rows, dQ / dK / dV stores).
#pragma once
#include <cstdint>

namespace sam3b {

constexpr int STG_BYTES = 2048;   // staging tile per warp

__device__ __forceinline__ uint32_t stg_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

__device__ __forceinline__ void stage_put_row(uint8_t* stg, int lane, const uint32_t (&w)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(stg + stg_off(lane, c)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}
__device__ __forceinline__ void stage_get_row(const uint8_t* stg, int lane, uint32_t (&w)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(stg + stg_off(lane, c));
    w[4 * c] = u.x; w[4 * c + 1] = u.y; w[4 * c + 2] = u.z; w[4 * c + 3] = u.w;
  }
}
// staging -> global.  g = address of (first row of the warp, first byte of the segment); rows_valid / bytes_valid clip.
__device__ __forceinline__ void stage_flush(const uint8_t* stg, int lane, uint8_t* g, int64_t ld_bytes, int rows_valid, int bytes_valid) {
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2);
    const uint4 u = *reinterpret_cast<const uint4*>(stg + stg_off(row, c));
    if (row < rows_valid && c * 16 < bytes_valid) *reinterpret_cast<uint4*>(g + (int64_t)row * ld_bytes + c * 16) = u;
  }
}
// global -> staging (same access shape)
__device__ __forceinline__ void stage_fill(uint8_t* stg, int lane, const uint8_t* g, int64_t ld_bytes, int rows_valid, int bytes_valid) {
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2);
    uint4 u = make_uint4(0, 0, 0, 0);
    if (row < rows_valid && c * 16 < bytes_valid) u = *reinterpret_cast<const uint4*>(g + (int64_t)row * ld_bytes + c * 16);
    *reinterpret_cast<uint4*>(stg + stg_off(row, c)) = u;
  }
}

// Asynchronous form of stage_fill_rows: cp.async (LDGSTS) copies global -> staging without passing through registers, so
// the caller can issue it, wait on an mbarrier, and only then stage_async_wait() + __syncwarp() before reading the tile:
// the global latency overlaps the wait and costs no registers (a register prefetch here spilled, see attn_bwd.cu).
template <typename RowFn>
__device__ __forceinline__ void stage_fill_rows_async(uint8_t* stg, int lane, RowFn rowaddr) {
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2);
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(stg + stg_off(row, c)));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rowaddr(row) + c * 16) : "memory");
  }
}
__device__ __forceinline__ void stage_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// global -> staging with one address per row: rowaddr(r) = first byte of row r's 64-byte segment (always readable)
template <typename RowFn>
__device__ __forceinline__ void stage_fill_rows(uint8_t* stg, int lane, RowFn rowaddr) {
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(rowaddr(row)) + c);
    *reinterpret_cast<uint4*>(stg + stg_off(row, c)) = u;
  }
}

}  // namespace sam3b
