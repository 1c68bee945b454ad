// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M][N] = epilogue( alpha * A[M][K] . B[N][K]^T )
//
// This one kernel carries every dense contraction of the SAM3 ViT trunk (qkv / proj / fc1 /
// fc2 forward, their dgrads, the skinny LoRA down-projections and the LoRA weight-gradients):
// reference call sites sam3/model/vitdet.py:480,513 (qkv, proj), timm Mlp fc1/fc2
// (vitdet.py:585-590), lora_layers.py:54-55,91 (x@A@B*alpha/r added to the base Linear).
// The LoRA up-projection never runs as its own Linear: the caller appends s*(x.A) as extra K
// columns of A and lora_B as extra K columns of the weight, so it rides the same K loop.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one elected lane, tcgen05.mma cta_group::1 kind::f16, M=128,N=BN,K=16)
//   warp 2   : TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-11: epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), and a static
// persistent tile schedule (tile = blockIdx.x + i*gridDim.x, n fastest so CTAs running
// together share the A row-panel in L2).
#include "gemm.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"
#include "gelu.cuh"
#include "stage.cuh"
#include "rng.cuh"

namespace sam3b {

struct GemmParams {
  int M, N, K, splitk;
  void* C; int64_t ldc;
  void* C2; int64_t ldc2;
  const float* bias;
  const float* res; int64_t ldres; int res_row_mod;
  const float* row_scale; int rows_per_scale;
  float drop_inv_keep; uint32_t drop_seed, drop_thr; const uint32_t* drop_seed_dev;
  const void* aux; int64_t ldaux;
  const float2* rope; int rope_period; int rope_cols;
  float alpha;
  int c_trans;
  uint32_t mn_lbo, mn_sbo;
  float* delta; int delta_Lq, delta_Lq_stat; int64_t delta_stride;
  int tma_store;   // gemm2_kernel: bit 0 = C, bit 1 = C2 go out through TMA stores (16-bit outputs); bit 2 = aux comes in by TMA
};


template <int BN>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TCOLS = 2 * BN;  // 512 / 256 / 128: powers of two >= 32
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + 8 * STG_BYTES + 1024;
};

// ---------------------------------------------------------------------------------------
// Coalesced epilogue I/O.  tcgen05.ld 32x32b hands every lane one output ROW, so a direct 16-byte store per lane
// touches 32 different 128-byte lines per instruction (32 LSU wavefronts, half-filled sectors): with two 16-bit outputs
// the GELU epilogue kept the L1->XBAR path ~60 % busy and the epilogue, not the MMA, paced the kernel
// (profiles/r01_ncu_gemm2_kernelILi3.txt).  Each epilogue warp therefore owns a 2 KB staging tile in shared memory
// (32 rows x 64 bytes): lanes write / read their own row segment, and global memory is accessed with 4 lanes per row, so
// one instruction covers 8 rows x 64 contiguous bytes (16 full sectors).  The 16-byte chunks are XOR-swizzled by
// (row >> 1) & 3, which is conflict-free for both access patterns.
// ---------------------------------------------------------------------------------------

// Per-warp state of the TMA-store epilogue (gemm2_kernel).  The staging tile IS the TMA box (32 rows x 64 bytes, 64-byte
// swizzle = stage.cuh's chunk XOR), so after the lanes have written their rows one elected lane issues ONE
// cp.async.bulk.tensor per 32 x 32 block instead of the warp reading the tile back and issuing 8-row global stores: per 2 KB
// block the LSU data pipe sees 16 wavefronts (the st.shared) instead of ~60 (st.shared + ld.shared + 8-row STGs at half-line
// efficiency).  That pipe is shared with the tensor core's operand reads: ncu on the GELU GEMM showed 58 % LSU + 31 % tensor
// wavefronts = a saturated pipe with the tensor pipe only 66 % active (profiles/r02_ncu_gemm2_gelu.txt).
// tmc == nullptr: staged global stores (gemm_kernel, conv3x3_kernel, and fp32 outputs everywhere).
struct TmaStore {
  const CUtensorMap* tmc = nullptr;    // map of C
  const CUtensorMap* tmc2 = nullptr;   // map of C2 (GELU epilogue)
  uint8_t* tile0 = nullptr;            // two adjacent 512-byte-aligned staging tiles of this warp: tile0, tile0 + STG_BYTES
  int flip = 0;
  // aux operand (h of the GELU' epilogue, O of the delta epilogue) fetched by TMA into tile0, one block ahead: the kernel
  // issues the tile's first block before it waits for the accumulator, each block issues its successor once it has read tile0
  const CUtensorMap* tmaux = nullptr;
  uint64_t* auxbar = nullptr;          // this warp's mbarrier
  uint32_t aux_phase = 0;
  int aux_next_col = -1;               // column of the next block of this warp in this tile, -1: none
};

// Request the next aux block into tile0.  Call it only AFTER every lane has CONSUMED the words it read from tile0 (the
// ld.shared data has then arrived; a __syncwarp right after issuing the loads does not order them against the async-proxy
// write of the TMA unit: a full-size run showed run-to-run gradient differences of 3 % with the request placed there).
__device__ __forceinline__ void aux_prefetch_next(TmaStore* ts, int lane, int row0);
__device__ __forceinline__ void aux_issue(TmaStore* ts, int lane, int col, int row0) {
  if (lane == 0) {
    mbar_arrive_expect_tx(ts->auxbar, STG_BYTES);
    tma_load_2d(ts->tile0, ts->tmaux, ts->auxbar, col, row0);
  }
}
__device__ __forceinline__ void aux_prefetch_next(TmaStore* ts, int lane, int row0) {
  if (ts == nullptr || ts->tmaux == nullptr) return;
  __syncwarp();
  if (ts->aux_next_col >= 0) aux_issue(ts, lane, ts->aux_next_col, row0);
}
// 32 x 32 block of the 16-bit aux operand -> this lane's row (16 packed words); falls back to the staged global loads
template <typename FillFn>
__device__ __forceinline__ void aux_get_row(TmaStore* ts, uint8_t* stg, int lane, int row0, uint32_t (&w)[16], FillFn fill) {
  if (ts != nullptr && ts->tmaux != nullptr) {
    mbar_wait(ts->auxbar, ts->aux_phase, 500);
    ts->aux_phase ^= 1;
    stage_get_row(ts->tile0, lane, w);
  } else {
    fill();
    __syncwarp();
    stage_get_row(stg, lane, w);
    __syncwarp();
  }
}

// this lane's 32 values -> 16-bit -> one TMA store of rows [row0, row0+32) x columns [col, col+32).  OUTSTANDING = how many
// earlier stores of this warp may still be reading their tile (1 when the caller alternates tiles, 0 when it reuses one).
template <int DT, int OUTSTANDING>
__device__ __forceinline__ void store16x32_tma(uint8_t* tile, int lane, const CUtensorMap* tm, int row0, int col, const float (&v)[32]) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack2<DT>(v[2 * i], v[2 * i + 1]);
  if (lane == 0) tma_store_wait_read<OUTSTANDING>();
  __syncwarp();
  stage_put_row(tile, lane, w);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tm, tile, col, row0);
    tma_store_commit();
  }
}

// this lane's 32 values -> 16-bit -> rows [row0, row0+32) x columns [col, col+32) of `base` (leading dimension ld elements)
template <int DT>
__device__ __forceinline__ void store16x32(uint8_t* stg, int lane, void* base, int64_t ld, int row0, int col, const float (&v)[32],
                                           int rows_valid, int ncols_valid) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack2<DT>(v[2 * i], v[2 * i + 1]);
  stage_put_row(stg, lane, w);
  __syncwarp();
  stage_flush(stg, lane, reinterpret_cast<uint8_t*>(reinterpret_cast<uint16_t*>(base) + (int64_t)row0 * ld + col), ld * 2, rows_valid,
              ncols_valid * 2);
  __syncwarp();
}
__device__ __forceinline__ void store32x32(uint8_t* stg, int lane, float* base, int64_t ld, int row0, int col, const float (&v)[32],
                                           int rows_valid, int ncols_valid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {   // 16 fp32 columns = 64 bytes per pass
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[h * 16 + i]);
    stage_put_row(stg, lane, w);
    __syncwarp();
    stage_flush(stg, lane, reinterpret_cast<uint8_t*>(base + (int64_t)row0 * ld + col + h * 16), ld * 4, rows_valid,
                (ncols_valid - h * 16) * 4);
    __syncwarp();
  }
}
// rows [row0 + ..) x 32 fp32 columns of `base` -> this lane's row (v[i] += )
__device__ __forceinline__ void addload32x32(uint8_t* stg, int lane, const float* base, int64_t ld, int64_t grow0, bool contiguous_rows,
                                             int col, float (&v)[32], int rows_valid, int ncols_valid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    stage_fill(stg, lane, reinterpret_cast<const uint8_t*>(base + grow0 * ld + col + h * 16), ld * 4, rows_valid, (ncols_valid - h * 16) * 4);
    __syncwarp();
    uint32_t w[16];
    stage_get_row(stg, lane, w);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[h * 16 + i] += __uint_as_float(w[i]);
  }
}

// One epilogue step: the warp owns output rows [row0, row0+32) (this lane: row0 + lane), columns [col, col+32).
// Executed by all 32 lanes (the staged stores are cooperative); rows >= M are clipped by rows_valid.
template <int EPI, int DT, bool FULL>
__device__ __forceinline__ void epilogue_chunk_impl(const GemmParams& p, uint8_t* stg, int lane, int row0, int col, const uint32_t (&r)[32],
                                                    TmaStore* ts) {
  // 16-bit output block: TMA store when the kernel provides maps, staged global stores otherwise
  auto put16 = [&](void* base, int64_t ld, const CUtensorMap* tm, const float (&vals)[32], int which, int rows_valid_, int nvalid_) {
    if (ts != nullptr && tm != nullptr) {
      if (which < 0) {            // single-output epilogue without an aux load: alternate the two tiles
        store16x32_tma<DT, 1>(ts->tile0 + ts->flip * STG_BYTES, lane, tm, row0, col, vals);
        ts->flip ^= 1;
      } else if (which < 2) {     // two-output epilogue: output `which` owns tile `which`
        store16x32_tma<DT, 1>(ts->tile0 + which * STG_BYTES, lane, tm, row0, col, vals);
      } else {                    // epilogue whose aux load uses tile 0: stores go through tile 1 only
        store16x32_tma<DT, 0>(ts->tile0 + STG_BYTES, lane, tm, row0, col, vals);
      }
    } else {
      store16x32<DT>(stg, lane, base, ld, row0, col, vals, rows_valid_, nvalid_);
    }
  };
  const CUtensorMap* tmc = ts != nullptr ? ts->tmc : nullptr;
  const CUtensorMap* tmc2 = ts != nullptr ? ts->tmc2 : nullptr;
  const int nvalid = FULL ? 32 : min(32, p.N - col);  // multiple of 8 (host-checked); FULL folds every column predicate
  const int row = row0 + lane;
  const int rows_valid = min(32, p.M - row0);
  const bool row_ok = lane < rows_valid;
  float v[32];
  if (p.alpha == 1.f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
  }

  if constexpr (EPI != EPI_ATOMIC_F32 && EPI != EPI_DGELU && EPI != EPI_ADDMASK16) {
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q * 4 < nvalid) {
          float4 b = __ldg(b4 + q);
          v[q * 4] += b.x; v[q * 4 + 1] += b.y; v[q * 4 + 2] += b.z; v[q * 4 + 3] += b.w;
        }
      }
    }
  }

  if constexpr (EPI == EPI_STORE16) {
    put16(p.C, p.ldc, tmc, v, -1, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_STORE32) {
    if (p.row_scale != nullptr) {  // device-resident scale (1/grad-scale of the neck / mask-head backward, conv_ops.py)
      const float sc = __ldg(p.row_scale + min(row, p.M - 1) / p.rows_per_scale);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= sc;
    }
    store32x32(stg, lane, reinterpret_cast<float*>(p.C), p.ldc, row0, col, v, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    // 2-D axial RoPE on adjacent pairs (vitdet.py:68-90): (a,b) -> (a*cos - b*sin, a*sin + b*cos).
    // The table row is the token's position inside its rope period (window or image); the pair
    // index is (col % 64)/2 inside the 64-wide head. v (cols >= rope_cols) is stored unrotated.
    if (col < p.rope_cols) {
      const float4* t4 =
          reinterpret_cast<const float4*>(p.rope + (int64_t)(row % p.rope_period) * 32 + ((col & 63) >> 1));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 cs = __ldg(t4 + q);  // (cos0, sin0, cos1, sin1)
        float a0 = v[q * 4], b0 = v[q * 4 + 1], a1 = v[q * 4 + 2], b1 = v[q * 4 + 3];
        v[q * 4] = a0 * cs.x - b0 * cs.y;
        v[q * 4 + 1] = a0 * cs.y + b0 * cs.x;
        v[q * 4 + 2] = a1 * cs.z - b1 * cs.w;
        v[q * 4 + 3] = a1 * cs.w + b1 * cs.z;
      }
    }
    put16(p.C, p.ldc, tmc, v, -1, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_RESIDUAL_F32) {
    if (p.row_scale != nullptr) {  // stochastic depth: per-image 0 or 1/keep on the branch (vitdet.py:610-611)
      const float sc = __ldg(p.row_scale + min(row, p.M - 1) / p.rows_per_scale);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= sc;
    }
    // residual rows: row itself, or row % res_row_mod (positional table of the patch embedding).  A warp's 32 rows map to
    // 32 consecutive residual rows unless they wrap around the table, in which case each lane loads its own row.
    const int rr0 = p.res_row_mod > 0 ? (row0 % p.res_row_mod) : row0;
    if (p.res_row_mod > 0 && rr0 + 32 > p.res_row_mod) {
      const float4* r4 = reinterpret_cast<const float4*>(p.res + (int64_t)(row % p.res_row_mod) * p.ldres + col);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (row_ok && q * 4 < nvalid) {
          float4 x = r4[q];
          v[q * 4] += x.x; v[q * 4 + 1] += x.y; v[q * 4 + 2] += x.z; v[q * 4 + 3] += x.w;
        }
      }
    } else {
      addload32x32(stg, lane, p.res, p.ldres, rr0, true, col, v, rows_valid, nvalid);
    }
    store32x32(stg, lane, reinterpret_cast<float*>(p.C), p.ldc, row0, col, v, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_GELU) {
    put16(p.C, p.ldc, tmc, v, 0, rows_valid, nvalid);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    put16(p.C2, p.ldc2, tmc2, v, 1, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_DGELU) {
    // h (the fc1 pre-activation, 16-bit) through the staging tile: 64 bytes per row
    uint32_t hw[16];
    aux_get_row(ts, stg, lane, row0, hw, [&] {
      stage_fill(stg, lane, reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint16_t*>(p.aux) + (int64_t)row0 * p.ldaux + col),
                 p.ldaux * 2, rows_valid, nvalid * 2);
    });
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 hh = unpack2<DT>(hw[i]);
      v[2 * i] *= dgelu_erf(hh.x);
      v[2 * i + 1] *= dgelu_erf(hh.y);
    }
    aux_prefetch_next(ts, lane, row0);
    put16(p.C, p.ldc, tmc, v, 2, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_STORE16_DELTA) {
    // O (the forward attention output, 16-bit) through the staging tile, like h in the DGELU epilogue
    uint32_t ow[16];
    aux_get_row(ts, stg, lane, row0, ow, [&] {
      stage_fill(stg, lane, reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint16_t*>(p.aux) + (int64_t)row0 * p.ldaux + col),
                 p.ldaux * 2, rows_valid, nvalid * 2);
    });
    float dsum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 g = unpack2<DT>(pack2<DT>(v[2 * i], v[2 * i + 1]));   // the 16-bit dO the attention kernels will read
      const float2 o = unpack2<DT>(ow[i]);
      dsum = fmaf(g.x, o.x, dsum);
      dsum = fmaf(g.y, o.y, dsum);
    }
    // two 32-column chunks per 64-wide head: two commutative adds per (row, head) -> bitwise deterministic
    if (row_ok) atomicAdd(p.delta + (int64_t)(col >> 6) * p.delta_stride + (int64_t)(row / p.delta_Lq) * p.delta_Lq_stat + row % p.delta_Lq, dsum);
    aux_prefetch_next(ts, lane, row0);
    put16(p.C, p.ldc, tmc, v, 2, rows_valid, nvalid);
  } else if constexpr (EPI == EPI_ADDMASK16) {
    if (row_ok) {
    // data-gradient of the adapter branch under dropout: dx += mask/(1-p) * (dT''.A^T) [* gelu'(h) on the fc2 site]
    uint4* c4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.C) + (int64_t)row * p.ldc + col);
    const uint4* h4 = p.aux == nullptr ? nullptr
                                       : reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.aux) + (int64_t)row * p.ldaux + col);
    const uint32_t mask_seed = p.drop_seed + (p.drop_seed_dev != nullptr ? __ldg(p.drop_seed_dev) : 0u);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q * 8 < nvalid) {
        uint4 cu = c4[q];
        uint32_t cw[4] = {cu.x, cu.y, cu.z, cu.w};
        uint32_t hw[4] = {0, 0, 0, 0};
        if (h4 != nullptr) { const uint4 hu = h4[q]; hw[0] = hu.x; hw[1] = hu.y; hw[2] = hu.z; hw[3] = hu.w; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 c = unpack2<DT>(cw[e]);
          float a0 = v[q * 8 + 2 * e] * p.drop_inv_keep, a1 = v[q * 8 + 2 * e + 1] * p.drop_inv_keep;
          if (h4 != nullptr) { const float2 hh = unpack2<DT>(hw[e]); a0 *= dgelu_erf(hh.x); a1 *= dgelu_erf(hh.y); }
          if (dropout_keep(mask_seed, row, col + q * 8 + 2 * e, p.N, p.drop_thr)) c.x += a0;
          if (dropout_keep(mask_seed, row, col + q * 8 + 2 * e + 1, p.N, p.drop_thr)) c.y += a1;
          cw[e] = pack2<DT>(c.x, c.y);
        }
        c4[q] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      }
    }
    }
  } else if constexpr (EPI == EPI_ATOMIC_F32) {
    float* C = reinterpret_cast<float*>(p.C);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (row_ok && i < nvalid) {
        int64_t off = p.c_trans ? ((int64_t)(col + i) * p.ldc + row) : ((int64_t)row * p.ldc + col + i);
        atomicAdd(C + off, v[i]);
      }
    }
  }
}

template <int EPI, int DT>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, uint8_t* stg, int lane, int row0, int col, const uint32_t (&r)[32],
                                               TmaStore* ts = nullptr) {
  if (row0 >= p.M || col >= p.N) return;   // warp-uniform
  if (col + 32 <= p.N) epilogue_chunk_impl<EPI, DT, true>(p, stg, lane, row0, col, r, ts);
  else epilogue_chunk_impl<EPI, DT, false>(p, stg, lane, row0, col, r, ts);
}

template <int BN, int EPI, int DT, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(384, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stg = reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES + ((threadIdx.x >> 5) >= 4 ? ((threadIdx.x >> 5) - 4) * STG_BYTES : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_tiles = (p.M + 127) / 128;
  const int kb_total = (p.K + 63) / 64;
  const int kb_per_split = (kb_total + p.splitk - 1) / p.splitk;
  const int total_work = m_tiles * n_tiles * p.splitk;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.splitk;
        const int tile = w / p.splitk;
        const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
        const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
          mbar_arrive_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
          uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tmA, &full[stage], kb * 64, m_blk * 128);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d(a_dst + c * 8192, &tmA, &full[stage], m_blk * 128 + c * 64, kb * 64);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b_dst, &tmB, &full[stage], kb * 64, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * 8192, &tmB, &full[stage], n_blk * BN + c * 64, kb * 64);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, BN, DT, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.splitk;
        const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        mbar_wait(&tempty[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
          const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
          for (int k = 0; k < ksteps; ++k) {
            uint64_t da, db;
            if constexpr (!A_MN) da = make_desc_kmajor(a_addr + k * 32);
            else da = make_smem_desc_sw128(a_addr + k * 2048, p.mn_lbo, p.mn_sbo);
            if constexpr (!B_MN) db = make_desc_kmajor(b_addr + k * 32);
            else db = make_smem_desc_sw128(b_addr + k * 2048, p.mn_lbo, p.mn_sbo);
            umma_f16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // smem slot is free once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    // 8 epilogue warps: two per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31), each
    // owning one half of the tile's columns, so the epilogue math/stores of a tile finish well
    // inside the next tile's MMA time even for the GELU / GELU' epilogues.
    const int ew = warp & 3;
    const int c_lo = ((warp - 4) >> 2) * (BN / 2);
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w / p.splitk;
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      mbar_wait(&tfull[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const int row0 = m_blk * 128 + ew * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + BN / 2; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(t_row + c, r);
        tmem_ld_wait();
        epilogue_chunk<EPI, DT>(p, stg, lane, row0, n_blk * BN + c, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);   // TMEM hand-off: no need to wait for the stores above
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TCOLS);
}


// ---------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs owns one 256 x 256 output tile.  Each CTA
// loads its own 128 rows of A and its own 128-row half of the B tile (32 KB per 64-deep k-block
// instead of 48 KB), the leader issues M=256 MMAs that read both CTAs' shared memory, and each
// CTA drains its own 128 accumulator rows.  Per FLOP this moves 2/3 of the L2->SM bytes of the
// single-CTA 128x256 tile, which is what bounded the single-CTA kernel (~13 TB/s of L2 reads at
// 1.1 PFLOP/s).  K-major operands, BN = 256 only.
// ---------------------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int BM = 128, BN = 256, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;        // 16 KB (this CTA's rows)
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // 16 KB (this CTA's half of the N tile)
  static constexpr int STAGES = 6;
  static constexpr int TCOLS = 512;
  // Epilogue warps: 8 (2 per scheduler).  Round-1 A/B on B200: 16 warps (4 per scheduler, <= 96 registers) and a GELU
  // with a third fewer instructions both left the GELU / GELU' GEMMs at the same time relative to the plain-store GEMM
  // (+25 % / +37 %), i.e. the epilogue is neither latency- nor issue-bound; the extra time tracks the extra 393 MB of
  // HBM traffic per launch under the 1000 W power cap (SM clock 1.5-1.68 GHz of 1.965 during these runs).
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);
  // barriers (256 B, padded to 512 so the staging tiles are 512-byte aligned: the TMA 64-byte swizzle is a function of the
  // absolute shared-memory address bits [7:8]) + two staging tiles per epilogue warp
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + 2 * EPI_WARPS * STG_BYTES + 1024;
  static_assert(SMEM <= 232448, "gemm2_kernel: shared memory over the 227 KB limit");
};

template <int EPI, int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg::THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
             const __grid_constant__ CUtensorMap tmC2, const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* full = bars;                 // used in the leader CTA only
  uint64_t* empty = bars + STAGES;       // in both CTAs (multicast commit)
  uint64_t* tfull = bars + 2 * STAGES;   // in both CTAs (multicast commit)
  uint64_t* tempty = tfull + 2;          // used in the leader CTA only, one arrival per epilogue warp of both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* auxbar = bars + 32;          // byte 256 of the barrier block: one mbarrier per epilogue warp (aux TMA loads)
  uint8_t* stg = reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES + ((threadIdx.x >> 5) >= 4 ? ((threadIdx.x >> 5) - 4) * 2 * STG_BYTES : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_pairs = (p.M + 255) / 256;
  const int kb_total = (p.K + 63) / 64;
  const int total_work = m_pairs * n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store & 1) tma_prefetch_desc(&tmC);
    if (p.tma_store & 6) tma_prefetch_desc(&tmC2);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * Cfg::EPI_WARPS); }
    for (int i = 0; i < Cfg::EPI_WARPS; ++i) mbar_init(&auxbar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TCOLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += n_clusters) {
        const int n_blk = w % n_tiles, m_pair = w / n_tiles;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
          const uint32_t full_leader = mapa_u32(&full[stage], 0);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          tma_load_2d_pair(sA + stage * Cfg::A_BYTES, &tmA, full_leader, kb * 64, m_pair * 256 + (int)rank * 128);
          tma_load_2d_pair(sB + stage * Cfg::B_BYTES, &tmB, full_leader, kb * 64, n_blk * BN + (int)rank * 128);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN, DT, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = cluster_id; w < total_work; w += n_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
          const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_f16_ss_pair(d_tmem, make_desc_kmajor(a_addr + k * 32), make_desc_kmajor(b_addr + k * 32), idesc,
                             (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty[stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    const int ew = warp & 3;
    const int c_lo = ((warp - 4) >> 2) * Cfg::COLS_PER_WARP;
    TmaStore ts_state;
    ts_state.tmc = (p.tma_store & 1) ? &tmC : nullptr;
    ts_state.tmc2 = (p.tma_store & 2) ? &tmC2 : nullptr;
    ts_state.tile0 = stg;
    ts_state.tmaux = (p.tma_store & 4) ? &tmC2 : nullptr;
    ts_state.auxbar = &auxbar[warp - 4];
    TmaStore* ts = p.tma_store ? &ts_state : nullptr;
    const bool aux_tma = (p.tma_store & 4) != 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = cluster_id; w < total_work; w += n_clusters) {
      const int n_blk = w % n_tiles, m_pair = w / n_tiles;
      const int row0 = m_pair * 256 + (int)rank * 128 + ew * 32;
      // the aux block of this warp's first chunk is requested before the accumulator wait (same skip rule as epilogue_chunk)
      if (aux_tma && row0 < p.M && n_blk * BN + c_lo < p.N) aux_issue(ts, lane, n_blk * BN + c_lo, row0);
      mbar_wait(&tfull[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + Cfg::COLS_PER_WARP; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(t_row + c, r);
        tmem_ld_wait();
        if (aux_tma) {
          const int nc = n_blk * BN + c + 32;
          ts_state.aux_next_col = (c + 32 < c_lo + Cfg::COLS_PER_WARP && nc < p.N) ? nc : -1;
        }
        epilogue_chunk<EPI, DT>(p, stg, lane, row0, n_blk * BN + c, r, ts);
      }
      tc_fence_before();
      __syncwarp();
      #ifdef SAM3B_TEMPTY_RELEASE   // A/B switch for the measurement in profiles/README.md
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
#else
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(&tempty[acc], 0));
#endif
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (ts != nullptr && lane == 0) tma_store_wait_all();   // this lane's bulk stores: sources read and global writes performed
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::TCOLS);
}

template <int EPI, int DT>
static int launch_pair(const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg;
  CUtensorMap tmA, tmB, tmC, tmC2;
  int rc;
  if ((rc = make_tmap_2d(&tmA, a.A, a.M, a.K, a.lda, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tmB, a.B, a.N, a.K, a.ldb, 128, 64))) return rc;
  // 16-bit outputs leave through TMA stores (SAM3B_TMA_STORE=0: staged global stores, for A/B measurements)
  static const bool tma_store_on = [] { const char* e = getenv("SAM3B_TMA_STORE"); return !(e && e[0] == '0'); }();
  GemmParams pp = p;
  pp.tma_store = 0;
  constexpr bool kStore16 = EPI == EPI_STORE16 || EPI == EPI_QKV_ROPE || EPI == EPI_GELU || EPI == EPI_DGELU || EPI == EPI_STORE16_DELTA;
  std::memset(&tmC, 0, sizeof(tmC));
  std::memset(&tmC2, 0, sizeof(tmC2));
  if (kStore16 && tma_store_on && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (a.ldc * 2) % 16 == 0) {
    if ((rc = make_tmap_store16(&tmC, a.C, a.M, a.N, a.ldc))) return rc;
    pp.tma_store |= 1;
    if (EPI == EPI_GELU) {
      if ((reinterpret_cast<uintptr_t>(a.C2) & 15) == 0 && (a.ldc2 * 2) % 16 == 0) {
        if ((rc = make_tmap_store16(&tmC2, a.C2, a.M, a.N, a.ldc2))) return rc;
        pp.tma_store |= 2;
      } else {
        pp.tma_store = 0;      // both outputs or neither: the two-tile scheme assumes both stores are TMA stores
      }
    }
  }
  // GELU' / delta epilogues: the 16-bit aux operand comes in through TMA loads into the warp's first staging tile (the C2 map
  // slot is free in these epilogues); needs the TMA stores on (tile 1 is then the only store tile)
  constexpr bool kAux = EPI == EPI_DGELU || EPI == EPI_STORE16_DELTA;
  static const bool tma_aux_on = [] { const char* e = getenv("SAM3B_TMA_AUX"); return !(e && e[0] == '0'); }();
  if (kAux && tma_aux_on && (pp.tma_store & 1) && a.aux != nullptr && (reinterpret_cast<uintptr_t>(a.aux) & 15) == 0 && (a.ldaux * 2) % 16 == 0) {
    if ((rc = make_tmap_store16(&tmC2, a.aux, a.M, a.N, a.ldaux))) return rc;
    pp.tma_store |= 4;
  }
  auto kern = gemm2_kernel<EPI, DT>;
  static bool attr_set = false;
  if (!attr_set) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int total = ((a.M + 255) / 256) * ((a.N + 255) / 256);
  int clusters = (a.max_ctas > 0 ? a.max_ctas : num_sms()) / 2;
  if (clusters > total) clusters = total;
  if (clusters < 1) clusters = 1;
  SAM3B_CHECK_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(Cfg::THREADS), Cfg::SMEM, stream, tmA, tmB, tmC, tmC2, pp));
  SAM3B_LAUNCHED();
  return 0;
}
template <int EPI>
static int launch_pair_dt(const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
  return a.dtype == 0 ? launch_pair<EPI, 0>(a, p, stream) : launch_pair<EPI, 1>(a, p, stream);
}

// ---------------------------------------------------------------------------------------
// Implicit-GEMM 3x3 convolution (see gemm.cuh).  Same warp roles and mbarrier ring as gemm_kernel<256>; an output
// tile is a 16 x 8 pixel patch of one image (128 GEMM rows), so the A operand of filter tap (ky, kx), channels
// [c0, c0+64) is the 4-D TMA box {c0, x0+kx-1, y0+ky-1, b} of extent {64, 8, 16, 1}: in shared memory that is 128 rows
// of 128 bytes with the 128B swizzle, i.e. exactly the K-major tile the MMA descriptor expects.  Borders cost nothing:
// out-of-range coordinates are zero-filled by the TMA unit.  K loop = 9 taps x C/64 blocks.
// ---------------------------------------------------------------------------------------
struct ConvParams {
  int H, W, C, Cout, tiles_x, tiles_y, n_tiles, total_work, kb_total, kb_per_tap;
  const float* bias;
  void* out; int64_t ldc;
};
constexpr int CONV_BH = 16, CONV_BW = 8;

template <bool OUT32, int DT>
__global__ void __launch_bounds__(384, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using Cfg = GemmCfg<256>;
  constexpr int STAGES = Cfg::STAGES, BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const int n_blk = w % p.n_tiles, t = w / p.n_tiles;
        const int b = t / tiles_per_img, ti = t % tiles_per_img;
        const int y0 = (ti / p.tiles_x) * CONV_BH, x0 = (ti % p.tiles_x) * CONV_BW;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          const int tap = kb / p.kb_per_tap, c0 = (kb % p.kb_per_tap) * 64;
          mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
          mbar_arrive_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_4d(sA + stage * Cfg::A_BYTES, &tmX, &full[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmW, &full[stage], kb * 64, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, BN, DT, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&full[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(d_tmem, make_desc_kmajor(a_addr + k * 32), make_desc_kmajor(b_addr + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // lane = tile row ew*32 + lane = pixel (y0 + row / 8, x0 + row % 8); each lane owns whole 64/128-byte output segments
    const int ew = warp & 3;
    const int c_lo = ((warp - 4) >> 2) * (BN / 2);
    const int row = ew * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const int n_blk = w % p.n_tiles, t = w / p.n_tiles;
      const int b = t / tiles_per_img, ti = t % tiles_per_img;
      const int y = (ti / p.tiles_x) * CONV_BH + row / CONV_BW, x = (ti % p.tiles_x) * CONV_BW + row % CONV_BW;
      const bool ok = y < p.H;
      const int64_t m = ((int64_t)b * p.H + y) * p.W + x;
      mbar_wait(&tfull[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + BN / 2; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(t_row + c, r);
        tmem_ld_wait();
        const int col = n_blk * BN + c;
        if (!ok || col >= p.Cout) continue;
        const int nvalid = min(32, p.Cout - col);   // multiple of 8 (host-checked)
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q * 4 < nvalid) { const float4 bb = __ldg(b4 + q); v[q * 4] += bb.x; v[q * 4 + 1] += bb.y; v[q * 4 + 2] += bb.z; v[q * 4 + 3] += bb.w; }
        }
        if constexpr (OUT32) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldc + col);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q * 4 < nvalid) o[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + m * p.ldc + col);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q * 8 < nvalid)
              o[q] = make_uint4(pack2<DT>(v[q * 8], v[q * 8 + 1]), pack2<DT>(v[q * 8 + 2], v[q * 8 + 3]),
                                pack2<DT>(v[q * 8 + 4], v[q * 8 + 5]), pack2<DT>(v[q * 8 + 6], v[q * 8 + 7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);   // TMEM hand-off: no need to wait for the stores above
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TCOLS);
}

bool conv3x3_supported(int H, int W, int C, int Cout) { return H > 0 && W > 0 && W % CONV_BW == 0 && C > 0 && C % 64 == 0 && Cout > 0 && Cout % 8 == 0; }

template <bool OUT32, int DT>
static int launch_conv(const ConvArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<256>;
  CUtensorMap tmX, tmW;
  int rc;
  if ((rc = make_tmap_nhwc(&tmX, a.x16, a.B, a.H, a.W, a.C, 64, CONV_BW, CONV_BH))) return rc;
  if ((rc = make_tmap_2d(&tmW, a.w9, a.Cout, 9 * (uint64_t)a.C, 9 * (uint64_t)a.C, 256, 64))) return rc;
  ConvParams p{};
  p.H = a.H; p.W = a.W; p.C = a.C; p.Cout = a.Cout;
  p.tiles_x = a.W / CONV_BW; p.tiles_y = (a.H + CONV_BH - 1) / CONV_BH; p.n_tiles = (a.Cout + 255) / 256;
  p.total_work = a.B * p.tiles_x * p.tiles_y * p.n_tiles;
  p.kb_per_tap = a.C / 64; p.kb_total = 9 * p.kb_per_tap;
  p.bias = a.bias; p.out = a.out; p.ldc = a.ldc;
  auto kern = conv3x3_kernel<OUT32, DT>;
  static bool attr_set = false;
  if (!attr_set) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int ctas = std::min(num_sms(), p.total_work);
  kern<<<ctas, 384, Cfg::SMEM, stream>>>(tmX, tmW, p);
  SAM3B_LAUNCHED();
  return 0;
}

int conv3x3_launch(const ConvArgs& a, cudaStream_t stream) {
  SAM3B_REQUIRE(a.B > 0 && a.x16 && a.w9 && a.out, "conv3x3: null tensor / empty batch");
  SAM3B_REQUIRE(conv3x3_supported(a.H, a.W, a.C, a.Cout), "conv3x3: needs C %% 64 == 0, W %% 8 == 0, Cout %% 8 == 0 (H=%d W=%d C=%d Cout=%d)", a.H, a.W, a.C, a.Cout);
  SAM3B_REQUIRE(a.dtype == 0 || a.dtype == 1, "conv3x3: dtype %d (0 fp16, 1 bf16)", a.dtype);
  SAM3B_REQUIRE(a.ldc >= a.Cout && a.ldc % (a.out_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0, "conv3x3: output leading dimension / alignment");
  if (a.bias) SAM3B_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "conv3x3: bias not 16-byte aligned");
  if (a.out_f32) return a.dtype == 0 ? launch_conv<true, 0>(a, stream) : launch_conv<true, 1>(a, stream);
  return a.dtype == 0 ? launch_conv<false, 0>(a, stream) : launch_conv<false, 1>(a, stream);
}

// ---------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------
template <int BN, int EPI, int DT, bool A_MN, bool B_MN>
static int launch_one(const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap_2d(&tmA, a.A, a.M, a.K, a.lda, 128, 64);
  else rc = make_tmap_2d(&tmA, a.A, a.K, a.M, a.lda, 64, 64);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d(&tmB, a.B, a.N, a.K, a.ldb, BN, 64);
  else rc = make_tmap_2d(&tmB, a.B, a.K, a.N, a.ldb, 64, 64);
  if (rc) return rc;
  auto kern = gemm_kernel<BN, EPI, DT, A_MN, B_MN>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int m_tiles = (a.M + 127) / 128, n_tiles = (a.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * p.splitk;
  int ctas = a.max_ctas > 0 ? a.max_ctas : num_sms();
  if (ctas > total) ctas = total;
  SAM3B_CHECK_CUDA(launch_pdl(kern, dim3(ctas), dim3(384), Cfg::SMEM, stream, tmA, tmB, p));
  SAM3B_LAUNCHED();
  return 0;
}

template <int BN, int EPI, bool A_MN, bool B_MN>
static int launch_dt(const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
  if (a.dtype == 0) return launch_one<BN, EPI, 0, A_MN, B_MN>(a, p, stream);
  return launch_one<BN, EPI, 1, A_MN, B_MN>(a, p, stream);
}

// CTA-pair tiles are the default for large problems (validated on B200: same results, +8-10 %);
// SAM3B_GEMM_PAIR=0 (or GemmArgs::cta_pair = 1) selects the single-CTA kernel for A/B runs.
static bool pair_default() {
  static const bool on = [] {
    const char* e = getenv("SAM3B_GEMM_PAIR");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
  SAM3B_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  SAM3B_REQUIRE(a.N % 8 == 0, "gemm: N=%d must be a multiple of 8", a.N);
  // K needs no alignment: TMA zero-fills out-of-bounds columns/rows and the last k-block issues
  // ceil(K_rem/16) MMAs over them.
  SAM3B_REQUIRE(a.dtype == 0 || a.dtype == 1, "gemm: dtype %d (0 fp16, 1 bf16)", a.dtype);
  SAM3B_REQUIRE(a.A && a.B && a.C, "gemm: null operand");
  SAM3B_REQUIRE(a.a_mn == a.b_mn, "gemm: mixed operand majors are not instantiated");
  const int kb_total = (a.K + 63) / 64;
  int splitk = a.splitk < 1 ? 1 : a.splitk;
  if (splitk > kb_total) splitk = kb_total;
  // every split must own at least one k-block
  while (splitk > 1 && ((kb_total + splitk - 1) / splitk) * (splitk - 1) >= kb_total) --splitk;
  SAM3B_REQUIRE(splitk == 1 || a.epilogue == EPI_ATOMIC_F32, "gemm: split-K needs the atomic epilogue");

  GemmParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K; p.splitk = splitk;
  p.C = a.C; p.ldc = a.ldc; p.C2 = a.C2; p.ldc2 = a.ldc2;
  p.bias = a.bias;
  p.res = a.residual; p.ldres = a.ldres; p.res_row_mod = a.res_row_mod;
  p.row_scale = a.row_scale; p.rows_per_scale = a.rows_per_scale > 0 ? a.rows_per_scale : 1;
  p.drop_inv_keep = 1.f / (1.f - a.drop_p); p.drop_seed = a.drop_seed; p.drop_thr = dropout_threshold(a.drop_p); p.drop_seed_dev = a.drop_seed_dev;
  p.aux = a.aux; p.ldaux = a.ldaux;
  p.rope = reinterpret_cast<const float2*>(a.rope); p.rope_period = a.rope_period > 0 ? a.rope_period : 1;
  p.rope_cols = a.rope_cols;
  p.alpha = a.alpha; p.c_trans = a.c_trans;
  p.tma_store = 0;
  p.delta = a.delta; p.delta_Lq = a.delta_Lq > 0 ? a.delta_Lq : 1; p.delta_Lq_stat = a.delta_Lq_stat; p.delta_stride = a.delta_stride;
  p.mn_lbo = a.dbg_lbo > 0 ? a.dbg_lbo : 8192;
  p.mn_sbo = a.dbg_sbo > 0 ? a.dbg_sbo : 1024;

  const bool is16 = (a.epilogue == EPI_STORE16 || a.epilogue == EPI_QKV_ROPE || a.epilogue == EPI_GELU ||
                     a.epilogue == EPI_DGELU || a.epilogue == EPI_ADDMASK16 || a.epilogue == EPI_STORE16_DELTA);
  if (a.epilogue != EPI_ATOMIC_F32) {
    SAM3B_REQUIRE(a.ldc % (is16 ? 8 : 4) == 0, "gemm: ldc=%lld breaks 16-byte store alignment", (long long)a.ldc);
    SAM3B_REQUIRE((reinterpret_cast<uintptr_t>(a.C) & 15) == 0, "gemm: C not 16-byte aligned");
  }
  if (a.epilogue == EPI_QKV_ROPE) {
    SAM3B_REQUIRE(a.rope != nullptr && a.rope_cols % 64 == 0, "gemm: rope epilogue needs a table and 64-aligned rope_cols");
  }
  if (a.epilogue == EPI_RESIDUAL_F32) SAM3B_REQUIRE(a.residual != nullptr && a.ldres % 4 == 0, "gemm: residual epilogue needs residual with ld %% 4 == 0");
  if (a.epilogue == EPI_GELU) SAM3B_REQUIRE(a.C2 != nullptr && a.ldc2 % 8 == 0, "gemm: gelu epilogue needs C2");
  if (a.epilogue == EPI_DGELU) SAM3B_REQUIRE(a.aux != nullptr && a.ldaux % 8 == 0, "gemm: dgelu epilogue needs aux");
  if (a.epilogue == EPI_STORE16_DELTA)
    SAM3B_REQUIRE(a.aux != nullptr && a.ldaux % 8 == 0 && a.delta != nullptr && a.N % 64 == 0 && a.delta_Lq > 0 && a.bias == nullptr,
                  "gemm: delta epilogue needs aux (O), a zeroed delta buffer, N %% 64 == 0 and no bias");
  if (a.bias) SAM3B_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm: bias not 16-byte aligned");

  if (a.a_mn) {
    SAM3B_REQUIRE(a.epilogue == EPI_ATOMIC_F32, "gemm: MN-major operands are instantiated for the atomic epilogue only");
    return launch_dt<64, EPI_ATOMIC_F32, true, true>(a, p, stream);
  }
  int bn = a.bn;
  if (bn == 0) bn = (a.N <= 64) ? 64 : 256;
  if (bn == 64) {
    switch (a.epilogue) {
      case EPI_STORE16: return launch_dt<64, EPI_STORE16, false, false>(a, p, stream);
      case EPI_STORE32: return launch_dt<64, EPI_STORE32, false, false>(a, p, stream);
      case EPI_ATOMIC_F32: return launch_dt<64, EPI_ATOMIC_F32, false, false>(a, p, stream);
      default: return fail(-1, "gemm: epilogue %d not instantiated for BN=64", a.epilogue);
    }
  }
  SAM3B_REQUIRE(bn == 256, "gemm: bn must be 64 or 256 (got %d)", bn);
  const bool pair = a.cta_pair == 2 || (a.cta_pair == 0 && pair_default() && a.M >= 512 && a.N >= 256);
  if (pair) {
    switch (a.epilogue) {
      case EPI_STORE16: return launch_pair_dt<EPI_STORE16>(a, p, stream);
      case EPI_QKV_ROPE: return launch_pair_dt<EPI_QKV_ROPE>(a, p, stream);
      case EPI_RESIDUAL_F32: return launch_pair_dt<EPI_RESIDUAL_F32>(a, p, stream);
      case EPI_GELU: return launch_pair_dt<EPI_GELU>(a, p, stream);
      case EPI_DGELU: return launch_pair_dt<EPI_DGELU>(a, p, stream);
      case EPI_STORE16_DELTA: return launch_pair_dt<EPI_STORE16_DELTA>(a, p, stream);
      case EPI_STORE32: return launch_pair_dt<EPI_STORE32>(a, p, stream);
      case EPI_ADDMASK16: return launch_pair_dt<EPI_ADDMASK16>(a, p, stream);
      default: break;
    }
  }
  switch (a.epilogue) {
    case EPI_STORE16: return launch_dt<256, EPI_STORE16, false, false>(a, p, stream);
    case EPI_QKV_ROPE: return launch_dt<256, EPI_QKV_ROPE, false, false>(a, p, stream);
    case EPI_RESIDUAL_F32: return launch_dt<256, EPI_RESIDUAL_F32, false, false>(a, p, stream);
    case EPI_GELU: return launch_dt<256, EPI_GELU, false, false>(a, p, stream);
    case EPI_DGELU: return launch_dt<256, EPI_DGELU, false, false>(a, p, stream);
    case EPI_STORE16_DELTA: return launch_dt<256, EPI_STORE16_DELTA, false, false>(a, p, stream);
    case EPI_STORE32: return launch_dt<256, EPI_STORE32, false, false>(a, p, stream);
    case EPI_ADDMASK16: return launch_dt<256, EPI_ADDMASK16, false, false>(a, p, stream);
    default: return fail(-1, "gemm: epilogue %d not instantiated for BN=256", a.epilogue);
  }
}

}  // namespace sam3b
