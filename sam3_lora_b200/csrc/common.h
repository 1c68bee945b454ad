// Host-side helpers shared by the launchers: error reporting for the C ABI (no exceptions
// cross the boundary; every entry point returns 0 or a negative code and the message is
// fetched with sam3b_last_error()), TMA tensor-map construction, device properties.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace sam3b {

// Records a message (thread-local) and returns `code` (negative).
int fail(int code, const char* fmt, ...);
const char* last_error_message();

#define SAM3B_CHECK_CUDA(expr)                                                                     \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::sam3b::fail(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                              \
  } while (0)

// after every kernel launch: check the launch and count it (sam3b_launch_count(), used by bench.py's gpu_launches)
#define SAM3B_LAUNCHED()                                   \
  do {                                                     \
    SAM3B_CHECK_CUDA(cudaGetLastError());                  \
    ::sam3b::count_launch();                               \
  } while (0)

#define SAM3B_REQUIRE(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return ::sam3b::fail(-1, __VA_ARGS__);           \
  } while (0)

// 2-D row-major tensor [rows][cols] of 16-bit (or 32-bit) elements, leading dimension ld
// (elements); box = [box_rows][box_cols]; SWIZZLE_128B; out-of-bounds reads are zero-filled.
// Returns 0 on success.
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes = 2);

// Store-side map for the GEMM epilogues: 16-bit [rows][cols] tensor, box [32 rows][32 cols] = 32 x 64 bytes, SWIZZLE_64B —
// the layout of the epilogue's staging tile (stage.cuh: 16-byte chunk index XOR (row >> 1) & 3), so a tile written by the
// epilogue lanes is stored with one cp.async.bulk.tensor; rows / columns past the tensor's extent are clipped by the TMA unit.
int make_tmap_store16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld);

// 4-D map over a channels-last 16-bit tensor [B][H][W][C]: dimensions (C, W, H, B), box (box_c, box_w, box_h, 1),
// SWIZZLE_128B (box_c * 2 bytes == 128).  Coordinates may be negative / past the end: those elements read as zero,
// which is exactly the zero padding of a convolution.
int make_tmap_nhwc(CUtensorMap* out, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t box_c,
                   uint32_t box_w, uint32_t box_h);

int num_sms();
void count_launch();
long long launch_count();

// Programmatic dependent launch (PDL): the kernel may start (run its prologue: barrier init, TMEM allocation, descriptor
// prefetch) while the previous kernel on the stream is still draining; it calls griddepcontrol.wait (ptx.cuh pdl_wait) before
// touching global memory, so the visible order is unchanged.  Works under stream capture (programmatic graph edges).
// Off by default: SAM3B_PDL=1 in the environment turns the attribute on.  Round-2 A/B inside one gpurun call (bench.py, 10
// steps, alternating): 153.8 / 152.4 ms with it, 151.7 / 151.8 ms without - the step is paced by the 1000 W power cap, not by
// launch gaps, and early-resident dependent CTAs do not help (profiles/r02_pdl_ab.txt).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace sam3b
