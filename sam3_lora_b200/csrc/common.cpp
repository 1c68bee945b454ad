#include <cstdlib>
#include "common.h"

#include <cstdarg>
#include <atomic>
#include <cstdio>
#include <cudaTypedefs.h>

namespace sam3b {

static thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error_message() { return g_err; }

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // The driver entry point is resolved through the runtime so the library does not link
    // libcuda directly (it only exists on the GPU box).
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes) {
  auto fn = encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-1, "TMA base pointer %p not 16-byte aligned", ptr);
  if ((ld * elem_bytes) % 16 != 0) return fail(-1, "TMA row pitch %llu B not a multiple of 16", (unsigned long long)(ld * elem_bytes));
  if (box_cols * elem_bytes > 128) return fail(-1, "TMA inner box %u B exceeds the 128-byte swizzle span", box_cols * elem_bytes);
  if (box_rows > 256) return fail(-1, "TMA box rows %u > 256", box_rows);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-3, "cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return 0;
}

int make_tmap_store16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  auto fn = encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-1, "TMA store base pointer %p not 16-byte aligned", ptr);
  if ((ld * 2) % 16 != 0) return fail(-1, "TMA store row pitch %llu B not a multiple of 16", (unsigned long long)(ld * 2));
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-3, "cuTensorMapEncodeTiled (store) failed with %d (rows=%llu cols=%llu ld=%llu)", (int)r, (unsigned long long)rows,
                (unsigned long long)cols, (unsigned long long)ld);
  return 0;
}

int make_tmap_nhwc(CUtensorMap* out, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t box_c,
                   uint32_t box_w, uint32_t box_h) {
  auto fn = encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-1, "TMA base pointer %p not 16-byte aligned", ptr);
  if ((C * 2) % 16 != 0) return fail(-1, "TMA pixel pitch %llu B not a multiple of 16", (unsigned long long)(C * 2));
  if (box_c * 2 != 128) return fail(-1, "TMA inner box %u B must equal the 128-byte swizzle span", box_c * 2);
  if (box_w > 256 || box_h > 256) return fail(-1, "TMA box %ux%u > 256", box_w, box_h);
  cuuint64_t gdim[4] = {C, W, H, B};
  cuuint64_t gstride[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-3, "cuTensorMapEncodeTiled (NHWC) failed with %d (B=%llu H=%llu W=%llu C=%llu)", (int)r, (unsigned long long)B,
                (unsigned long long)H, (unsigned long long)W, (unsigned long long)C);
  return 0;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
  }();
  return n;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("SAM3B_PDL");      // opt-in: round-2 A/B on B200 measured it 0.5-1 % SLOWER per step
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

}  // namespace sam3b
