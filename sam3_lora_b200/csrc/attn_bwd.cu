// FlashAttention-style backward for the SAM3 ViT attention (head_dim 64) on tcgen05 + TMA.
//
// Reference semantics: autograd of sam3/model/vitdet.py:485-502 (apply_rotary_enc on q,k then
// F.scaled_dot_product_attention).  Given dO, the forward's q,k (rotated), v, the log2-domain
// LSE and delta = rowsum(dO*O):
//     P  = exp2(S*c - lse2)            S = q k^T, c = head_dim^-0.5 * log2(e)
//     dV = P^T dO        dP = dO V^T   dS = P * (dP - delta)
//     dQ = scale * dS K  dK = scale * dS^T Q        then the inverse rotation R^T on dQ, dK
// so that dqkv is the gradient w.r.t. the *un-rotated* qkv projection output.
//
// Two deterministic kernels, no atomics (recompute S/dP in both, 7 matmuls instead of 5):
//   attn_bwd_dkdv : work item = (128-key tile, head, segment), loop over 64-query blocks.
//                   S^T = K Q^T and dP^T = V dO^T in TMEM (thread = key row), dV += P^T dO, dK += dS^T Q.
//   attn_bwd_dq   : work item = (128-query tile, head, segment), loop over 64-key blocks.
//                   S = Q K^T, dP = dO V^T (thread = query row), dQ += dS K.
// Every MMA takes its A operand FROM TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): the item's stationary
// tiles (K, V resp. Q, dO; 16-bit, two elements per column) are parked in TMEM once and P^T / dS^T / dS are
// written back to TMEM by the compute warps.  Only the streamed 64-row B tiles live in shared memory: an
// SS-MMA with N=64 needs 192 B/clk of smem operands against the SM's 128 B/clk, which is what bounded the first
// version of these kernels (profiles/r01_attn_bwd_timeline.md).
// One persistent CTA per SM owns all 512 TMEM columns: S / dP are double-buffered and two groups of 8 compute
// warps alternate blocks (group g takes the blocks with (running block index & 1) == g), so the tensor core
// works on block j+1, j+2 while a group is in the exp / pack phase of block j.  Warp 16 = TMA producer (4-stage
// ring), warps 17 / 18 = the two MMA-issuing threads (17 also allocates TMEM).  The CTA walks a static list of work items; the next item's
// stationary tiles are parked before the current item's epilogue so its first MMAs overlap the epilogue.
#include "attn.cuh"

#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"
#include "stage.cuh"

namespace sam3b {

#ifdef SAM3B_TRACE
// Debug timeline (tools/attn_trace.py): one chosen CTA stamps clock64() at its pipeline events.
__device__ unsigned long long g_attn_trace[16384];
#define TR_CTA 70
// stamps items 2 and 3 of the traced CTA (steady state of the persistent loop): item 3 uses slots + 4096
#define TRACE(slot) do { if (blockIdx.x == TR_CTA && (it == 2 || it == 3) && (slot) < 4096) g_attn_trace[(slot) + (it - 2) * 4096] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

namespace {

constexpr int HD = 64;
constexpr int BT = 128;  // rows owned by a CTA (keys in dkdv, queries in dq)
constexpr int BI = 64;   // inner block (queries in dkdv, keys in dq)
constexpr int I_BYTES = BI * HD * 2;   // 8 KB
constexpr int NSI = 8;                 // ring depth of the streamed tiles (even: a stage always serves the same group).  A stage is
                                       // released by the block's LAST MMA but needed again by an MMA issued two blocks early, so the
                                       // effective prefetch distance is NSI - 2 blocks; 4 stages left TMA latency exposed (~450 clk/block)
constexpr int TCOLS = 512;
constexpr int NTHREADS = 19 * 32;      // 16 compute warps (2 groups of 8) + TMA warp + 2 MMA-issuing warps
constexpr int WARP_TMA = 16, WARP_MMA = 17, WARP_MMA2 = 18;

struct BwdParams {
  int Lq, Lk, tiles, H, nseg;
  const void* q; int64_t ldq;
  const void* kv; int64_t ldkv;
  const void* dO; int64_t lddo;
  int q_col0, k_col0, v_col0, do_col0;
  const float* lse2;
  const float* delta;
  int Lq_stat; int64_t stat_stride;    // statistics layout [H][nseg * Lq_stat]
  void* dq; int64_t lddq; int dq_col0;
  void* dkv; int64_t lddkv; int dk_col0, dv_col0;
  const float2* rope; int rope_period;  // null: no inverse rotation (MultiheadAttention sites)
  float scale_log2, scale;
  // GEN only
  const float* bias; const uint8_t* kpm;
  float drop_inv_keep; uint32_t drop_thr, drop_seed; const uint32_t* drop_bits; const uint32_t* drop_bitsT; int bias_vec4; int bits_pitch_k, bits_pitch_q;
};

__device__ __forceinline__ uint32_t attn_drop_base(uint32_t seed, uint32_t bh) { return lowbias32(seed ^ (bh * 0x9E3779B1u + 0x85EBCA6Bu)); }
__device__ __forceinline__ bool attn_drop_keep(uint32_t base, uint32_t q, uint32_t k, uint32_t Lk, uint32_t thr) {
  return lowbias32(base ^ (q * Lk + k)) >= thr;
}

template <int DT>
__device__ __forceinline__ void pack16(const float (&v)[32], uint32_t (&o)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = pack2<DT>(v[2 * i], v[2 * i + 1]);
}

// ---- coalesced item transitions (stage.cuh) ------------------------------------------------------------------------
// A compute warp owns the 32 rows of its TMEM lane quarter.  Row-per-lane global accesses (the first version of park /
// epilogue) touch 32 different lines per instruction: per item that was ~2k LSU wavefronts for parking K/V and ~4k for
// the rope-table loads + dK/dV stores, i.e. 1.3-2.0k and 4.1-5.3k clk of an 18.4k-clk window item
// (profiles/r01_attn_bwd_timeline.md).  Through the warp's 2 KB staging tile four lanes share a row, so an instruction
// covers 8 rows x 64 B: a quarter of the wavefronts.

// 32 rows x 32 16-bit elements (64 B per row) of global memory -> 16 packed TMEM columns of each lane (A-operand layout)
template <typename RowFn>
__device__ __forceinline__ void park_rows(uint8_t* stg, int lane, uint32_t taddr, RowFn rowaddr) {
  stage_fill_rows(stg, lane, rowaddr);
  __syncwarp();
  uint32_t v[16];
  stage_get_row(stg, lane, v);
  __syncwarp();
  tmem_st_x16(taddr, v);
}
// second half of an asynchronous park: the rows were requested with stage_fill_rows_async into stg
__device__ __forceinline__ void park_staged(uint8_t* stg, int lane, uint32_t taddr) {
  uint32_t v[16];
  stage_get_row(stg, lane, v);
  __syncwarp();
  tmem_st_x16(taddr, v);
}

// Inverse-RoPE table rows of the warp's 32 rows, columns [cc, cc+32) of the head: 16 (cos, sin) pairs = 128 B per row,
// requested asynchronously as two 64-byte halves into tiles srope and srope + STG_BYTES.
__device__ __forceinline__ void rope_fill_async(const BwdParams& p, uint8_t* srope, int lane, int64_t row0, int cc) {
  stage_fill_rows_async(srope, lane, [&](int rl) { return reinterpret_cast<const uint8_t*>(p.rope + (int64_t)((row0 + rl) % p.rope_period) * 32 + (cc >> 1)); });
  stage_fill_rows_async(srope + STG_BYTES, lane, [&](int rl) { return reinterpret_cast<const uint8_t*>(p.rope + (int64_t)((row0 + rl) % p.rope_period) * 32 + (cc >> 1) + 8); });
}

// out[row0 + lane][col .. col + 32) = (optionally inverse-rotated) acc * mul in 16-bit, for the warp's 32 rows.
// ROPE: the table halves were staged by the caller in tiles srope and srope + STG_BYTES (rope_fill_async, waited).
template <int DT, bool ROPE>
__device__ __forceinline__ void store_grad_rows(const BwdParams& p, uint8_t* stg, const uint8_t* srope, int lane, uint32_t taddr, void* out, int64_t ld,
                                                int64_t row0, int col, float mul, int rows_valid) {
  uint32_t t[32];
  tmem_ld_x32(taddr, t);
  tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t[i]) * mul;
  if constexpr (ROPE) {
    if (p.rope != nullptr) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t w[16];
        stage_get_row(srope + h * STG_BYTES, lane, w);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float cs = __uint_as_float(w[2 * q]), sn = __uint_as_float(w[2 * q + 1]);
          const float a = v[16 * h + 2 * q], b = v[16 * h + 2 * q + 1];
          v[16 * h + 2 * q] = a * cs + b * sn;
          v[16 * h + 2 * q + 1] = -a * sn + b * cs;
        }
      }
    }
  }
  uint32_t w[16];
  pack16<DT>(v, w);
  stage_put_row(stg, lane, w);
  __syncwarp();
  stage_flush(stg, lane, reinterpret_cast<uint8_t*>(reinterpret_cast<uint16_t*>(out) + row0 * ld + col), ld * 2, rows_valid, 64);
  __syncwarp();
}

// out[row][col0 + cc .. col0 + cc + 32) = (optionally inverse-rotated) acc * mul, 16-bit
template <int DT, bool ROPE>
__device__ __forceinline__ void store_grad_chunk(const BwdParams& p, void* out, int64_t ld, uint32_t taddr, int row, int rope_row,
                                                 int col0, int cc, float mul, bool valid) {
  uint32_t t[32];
  tmem_ld_x32(taddr + cc, t);
  tmem_ld_wait();
  if (!valid) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t[i]) * mul;
  if constexpr (ROPE) {
    if (p.rope != nullptr) {
      const float4* t4 = reinterpret_cast<const float4*>(p.rope + (int64_t)(rope_row % p.rope_period) * 32 + (cc >> 1));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 cs = __ldg(t4 + q);
        float a0 = v[q * 4], b0 = v[q * 4 + 1], a1 = v[q * 4 + 2], b1 = v[q * 4 + 3];
        v[q * 4] = a0 * cs.x + b0 * cs.y;
        v[q * 4 + 1] = -a0 * cs.y + b0 * cs.x;
        v[q * 4 + 2] = a1 * cs.z + b1 * cs.w;
        v[q * 4 + 3] = -a1 * cs.w + b1 * cs.z;
      }
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + (int64_t)row * ld + col0 + cc);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack2<DT>(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack2<DT>(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack2<DT>(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack2<DT>(v[q * 8 + 6], v[q * 8 + 7]);
    dst[q] = u;
  }
}

// 16-column variant: the item epilogue is spread over all 16 compute warps (4 column quarters x 4 lane quarters)
template <int DT, bool ROPE>
__device__ __forceinline__ void store_grad_chunk16(const BwdParams& p, void* out, int64_t ld, uint32_t taddr, int row, int rope_row,
                                                   int col0, int cc, float mul, bool valid) {
  uint32_t t[16];
  tmem_ld_x16(taddr + cc, t);
  tmem_ld_wait();
  if (!valid) return;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(t[i]) * mul;
  if constexpr (ROPE) {
    if (p.rope != nullptr) {
      const float4* t4 = reinterpret_cast<const float4*>(p.rope + (int64_t)(rope_row % p.rope_period) * 32 + (cc >> 1));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 cs = __ldg(t4 + q);
        float a0 = v[q * 4], b0 = v[q * 4 + 1], a1 = v[q * 4 + 2], b1 = v[q * 4 + 3];
        v[q * 4] = a0 * cs.x + b0 * cs.y;
        v[q * 4 + 1] = -a0 * cs.y + b0 * cs.x;
        v[q * 4 + 2] = a1 * cs.z + b1 * cs.w;
        v[q * 4 + 3] = -a1 * cs.w + b1 * cs.z;
      }
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + (int64_t)row * ld + col0 + cc);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 u;
    u.x = pack2<DT>(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack2<DT>(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack2<DT>(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack2<DT>(v[q * 8 + 6], v[q * 8 + 7]);
    dst[q] = u;
  }
}

// =========================================================================================
// dK / dV
// =========================================================================================
// TMEM columns: S^T x2 [0,128)  dP^T x2 [128,256)  dV [256,320)  dK [320,384)  K [384,416)  V [416,448)
//               P^T [448,480)  dS^T [480,512)   (K, V, P^T, dS^T: 16-bit A operands, two elements per column)
constexpr int DKDV_MAIN = 2 * NSI * I_BYTES + NSI * 2 * BI * 4 + 256;  // Q ring | dO ring | [lse2 | delta] ring | barriers
constexpr int DKDV_SMEM = DKDV_MAIN + 16 * STG_BYTES + 8 * 2 * STG_BYTES;   // + a row tile per compute warp, + 2 rope tiles per dK warp (199 KB)

// DROPB (GEN = false): the ViT instantiation + dropout through precomputed keep-bits, full blocks only (see attn_fwd.cu)
template <int DT, bool GEN, bool DROPB = false>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmI,   // q buffer, box [64][64]
                     const __grid_constant__ CUtensorMap tmdO,  // dO,       box [64][64]
                     const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sQ = smem_raw;                 // NSI x 8 KB
  uint8_t* sdO = sQ + NSI * I_BYTES;      // NSI x 8 KB
  float* sStat = reinterpret_cast<float*>(sdO + NSI * I_BYTES);  // [NSI][lse2 | delta][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + NSI * 2 * BI);
  uint64_t* in_full = bars;               // [NSI] TMA -> MMA, compute (statistics)
  uint64_t* in_free = in_full + NSI;      // [NSI] MMA commit -> TMA
  uint64_t* sdp_full = in_free + NSI;     // [2]   S^T/dP^T buffer b holds a new block
  uint64_t* sdp_free = sdp_full + 2;      // [2]   group b has it in registers
  uint64_t* pds_full = sdp_free + 2;      // [2]   group b has written P^T/dS^T
  uint64_t* acc_done = pds_full + 2;      // [2]   dV/dK MMAs of a block of group b complete (P^T/dS^T free again)
  uint64_t* kv_ready = acc_done + 2;      //       K, V of the item parked in TMEM
  uint64_t* kv_free = kv_ready + 1;       //       last S^T/dP^T MMA of the item complete
  uint64_t* all_done = kv_free + 1;       //       last dV/dK MMA of the item complete
  uint64_t* epi_done = all_done + 1;      //       accumulators of the item read back
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 1);

  const int warp = threadIdx.x >> 5;
  const int n_blocks = (p.Lq + BI - 1) / BI;     // 64-query blocks per item
  const int n_items = p.tiles * p.nseg * p.H;    // item = tile + tiles * (seg + nseg * head)

  if (warp == WARP_TMA && elect_one()) {
    tma_prefetch_desc(&tmI); tma_prefetch_desc(&tmdO);
    for (int i = 0; i < NSI; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&sdp_free[i], 256); mbar_init(&pds_full[i], 256); mbar_init(&acc_done[i], 1); }
    mbar_init(kv_ready, 512); mbar_init(kv_free, 1); mbar_init(all_done, 1); mbar_init(epi_done, 512);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 320;
  const uint32_t tm_K = tmem_base + 384, tm_V = tmem_base + 416, tm_P = tmem_base + 448, tm_dS = tmem_base + 480;

  if (warp == WARP_TMA) {
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int seg = (item / p.tiles) % p.nseg, head = item / (p.tiles * p.nseg);
        const int q_row0_seg = seg * p.Lq;
        for (int j = 0; j < n_blocks; ++j) {
          const int jb = it * n_blocks + j, st = jb % NSI;
          if (jb >= NSI) mbar_wait(&in_free[st], (jb / NSI - 1) & 1, 10);
          mbar_arrive_expect_tx(&in_full[st], 2 * I_BYTES + 2 * BI * 4);
          tma_load_2d(sQ + st * I_BYTES, &tmI, &in_full[st], p.q_col0 + head * HD, q_row0_seg + j * BI);
          tma_load_2d(sdO + st * I_BYTES, &tmdO, &in_full[st], p.do_col0 + head * HD, q_row0_seg + j * BI);
          // per-query statistics of this block (head-major layout: 64 consecutive floats each)
          const int64_t soff = (int64_t)head * p.stat_stride + (int64_t)seg * p.Lq_stat + j * BI;
          bulk_load_1d(sStat + st * (2 * BI), p.lse2 + soff, BI * 4, &in_full[st]);
          bulk_load_1d(sStat + st * (2 * BI) + BI, p.delta + soff, BI * 4, &in_full[st]);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // MMA stream 1: S^T / dP^T.  The issuing thread's scalar code (barrier polls, descriptor arithmetic) shares its
    // scheduler with four busy compute warps; with ONE thread issuing all 16 MMAs of a block that code took ~1050 clk
    // per block and was the kernel's critical path (profiles/r01_attn_bwd_timeline.md), hence two streams on two
    // sub-partitions and descriptors derived from per-kernel constants by one 64-bit add.
    if (elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_f16(BT, BI, DT, 0, 0);  // A (TMEM) K-major, B K-major, N=64
      constexpr uint32_t ST16 = I_BYTES >> 4;                          // one ring stage in descriptor address units
      const uint64_t dQk = make_desc_kmajor(smem_u32(sQ)), dOk = make_desc_kmajor(smem_u32(sdO));
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int B0 = it * n_blocks;
        mbar_wait(kv_ready, it & 1, 20);
        // S^T / dP^T of block j go out as soon as group (j&1) holds block j-2 in registers (sdp_free), i.e. they
        // overlap that group's exp / pack work; the other group is busy with block j-1 meanwhile.
        for (int j = 0; j < n_blocks; ++j) {
          const int jb = B0 + j, st = jb % NSI, b = jb & 1;
          mbar_wait(&in_full[st], (jb / NSI) & 1, 21);
          if (jb >= 2) mbar_wait(&sdp_free[b], ((jb >> 1) - 1) & 1, 24);
          tc_fence_after();
          const uint64_t q = dQk + (uint64_t)(st * ST16), o = dOk + (uint64_t)(st * ST16);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // S^T = K . Q^T
            umma_f16_ts(tm_S + b * 64, tm_K + k * 8, q + k * 2, idesc_kk, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dP^T = V . dO^T
            umma_f16_ts(tm_dP + b * 64, tm_V + k * 8, o + k * 2, idesc_kk, k > 0);
          umma_commit(&sdp_full[b]);
          if (j == n_blocks - 1) umma_commit(kv_free);
          if (j < 64) TRACE(1024 + (j & 63) * 4 + 0);
        }
      }
    }
  } else if (warp == WARP_MMA2) {
    // MMA stream 2: dV += P^T dO, dK += dS^T Q.  in_free may be signalled from here although the S^T/dP^T MMAs that
    // read the same stage belong to stream 1: pds_full(j) is only reached after sdp_full(j), i.e. after they completed.
    if (elect_one()) {
      constexpr uint32_t idesc_kmn = make_idesc_f16(BT, HD, DT, 0, 1); // A (TMEM) K-major, B MN-major, N=64
      constexpr uint32_t ST16 = I_BYTES >> 4;
      const uint64_t dQm = make_desc_mnmajor(smem_u32(sQ), 8192), dOm = make_desc_mnmajor(smem_u32(sdO), 8192);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int B0 = it * n_blocks;
        for (int j = 0; j < n_blocks; ++j) {
          const int jb = B0 + j, st = jb % NSI, b = jb & 1;
          mbar_wait(&pds_full[b], (jb >> 1) & 1, 22);
          if (j == 0 && it > 0) mbar_wait(epi_done, (it - 1) & 1, 25);  // previous item's dV/dK have been read back
          tc_fence_after();
          const uint64_t q = dQm + (uint64_t)(st * ST16), o = dOm + (uint64_t)(st * ST16);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dV += P^T . dO   (K = 64 queries; 16 rows = 2048 B per step)
            umma_f16_ts(tm_dV, tm_P + k * 8, o + k * 128, idesc_kmn, (j > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dK += dS^T . Q
            umma_f16_ts(tm_dK, tm_dS + k * 8, q + k * 128, idesc_kmn, (j > 0 || k > 0));
          umma_commit(&in_free[st]);
          umma_commit(&acc_done[b]);
          if (j == n_blocks - 1) umma_commit(all_done);
          if (j < 64) TRACE(1024 + (j & 63) * 4 + 1);
        }
      }
    }
  } else if (warp < 16) {
    const int g = warp >> 3;           // compute group
    const int wi = warp & 7;
    const int r = threadIdx.x & 127;   // key row inside the tile == TMEM lane
    const int cc = (wi >> 2) * 32;     // this warp's half of the 64 inner (query) columns
    const uint32_t lane_off = static_cast<uint32_t>((wi & 3) * 32) << 16;
    const float c = p.scale_log2;

    uint8_t* stg = smem_raw + DKDV_MAIN + warp * STG_BYTES;
    uint8_t* srope = smem_raw + DKDV_MAIN + 16 * STG_BYTES + wi * 2 * STG_BYTES;   // used by group 1 (dK) only
    const int lane = threadIdx.x & 31, lq = wi & 3;

    // K (group 0) / V (group 1) rows of an item -> TMEM; each warp parks 32 rows x 32 of the 64 elements
    auto park = [&](int item) {
      const int tile = item % p.tiles, seg = (item / p.tiles) % p.nseg, head = item / (p.tiles * p.nseg);
      const int col = (g == 0 ? p.k_col0 : p.v_col0) + head * HD + (wi >> 2) * 32;
      park_rows(stg, lane, (g == 0 ? tm_K : tm_V) + lane_off + (wi >> 2) * 16, [&](int rl) {
        const int64_t grow = (int64_t)seg * p.Lk + min(tile * BT + lq * 32 + rl, p.Lk - 1);
        return reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint16_t*>(p.kv) + grow * p.ldkv + col);
      });
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(kv_ready);
    };
    if ((int)blockIdx.x < n_items) park(blockIdx.x);

    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item % p.tiles, seg = (item / p.tiles) % p.nseg, head = item / (p.tiles * p.nseg);
      const int B0 = it * n_blocks;
      const int t_row0 = seg * p.Lk + tile * BT;   // this item's 128 keys
      const int k_in_seg = min(tile * BT + r, p.Lk - 1);
      bool key_masked = false;
      const float* bias_col = nullptr;
      uint32_t drop_base = 0;
      if constexpr (GEN) {
        if (p.kpm != nullptr) key_masked = __ldg(p.kpm + (int64_t)seg * p.Lk + k_in_seg) != 0;
        if (p.bias != nullptr) bias_col = p.bias + (int64_t)(seg * p.H + head) * p.Lq * p.Lk + k_in_seg;
        drop_base = attn_drop_base(p.drop_seed, (uint32_t)(seg * p.H + head));
      }
      const uint32_t* bitsT_row = nullptr;   // precomputed keep-bits of this key column (one word per 32 queries)
      if constexpr (GEN || DROPB) {
        if (p.drop_bitsT != nullptr) bitsT_row = p.drop_bitsT + ((int64_t)(seg * p.H + head) * p.Lk + k_in_seg) * p.bits_pitch_q;
      }
      for (int j = (B0 + g) & 1; j < n_blocks; j += 2) {   // blocks whose running index has parity g
        const int jb = B0 + j, st = jb % NSI;
        const float* stat = sStat + st * (2 * BI);
        const int q_valid = min(BI, p.Lq - j * BI);   // queries of this block that exist
        const bool tr = (threadIdx.x == 0 || threadIdx.x == 256) && j < 64;
        if (tr) TRACE(64 + j * 8 + 0);
        mbar_wait(&in_full[st], (jb / NSI) & 1, 33);     // statistics landed next to Q / dO
        mbar_wait(&sdp_full[g], (jb >> 1) & 1, 30);
        tc_fence_after();
        if (tr) TRACE(64 + j * 8 + 1);
        uint32_t s[32], d[32];
        tmem_ld_x32(tm_S + g * 64 + lane_off + cc, s);
        tmem_ld_x32(tm_dP + g * 64 + lane_off + cc, d);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sdp_free[g]);  // S^T/dP^T buffer g may be overwritten by block j+2
        if (tr) TRACE(64 + j * 8 + 2);
        float pv[32], dsv[32];
        // predicate-free path: full block, no additive mask, no key padding, and dropout either off or available as keep-bits
        // (the encoder's 5184 x 5184 self-attention with dropout 0.1 takes it; the general path below costs ~5x more)
        bool fast = q_valid == BI;
        if constexpr (GEN) fast = fast && bias_col == nullptr && p.kpm == nullptr && (p.drop_thr == 0 || bitsT_row != nullptr);
        if (fast) {
          uint32_t mwf = 0xffffffffu;
          float inv_keep = 1.f;
          if constexpr (GEN || DROPB) {
            if (DROPB || p.drop_thr != 0) { mwf = __ldg(bitsT_row + ((j * BI + cc) >> 5)); inv_keep = p.drop_inv_keep; }
          }
          const float4* l4 = reinterpret_cast<const float4*>(stat + cc);
          const float4* d4 = reinterpret_cast<const float4*>(stat + BI + cc);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 l = l4[q], dl = d4[q];
            const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = q * 4 + e;
              const float pe = ex2_approx(fmaf(__uint_as_float(s[i]), c, -lv[e]));
              if constexpr (GEN || DROPB) {
                const float ks = ((mwf >> i) & 1u) ? inv_keep : 0.f;
                pv[i] = pe * ks;
                dsv[i] = pe * fmaf(__uint_as_float(d[i]), ks, -dv[e]);
              } else {
                pv[i] = pe;
                dsv[i] = pe * (__uint_as_float(d[i]) - dv[e]);
              }
            }
          }
        } else {
          uint32_t mw = 0u;
          if constexpr (GEN) {
            if (bitsT_row != nullptr) { const int wi_ = (j * BI + cc) >> 5; mw = wi_ < p.bits_pitch_q ? __ldg(bitsT_row + wi_) : 0u; }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int qi = j * BI + cc + i;   // query index inside the segment
            float t = __uint_as_float(s[i]) * c;
            float dp = __uint_as_float(d[i]);
            bool dead = (cc + i) >= q_valid;
            float keep_scale = 1.f;
            if constexpr (GEN) {
              if (bias_col != nullptr && !dead) t = fmaf(__ldg(bias_col + (int64_t)qi * p.Lk), 1.4426950408889634f, t);
              dead = dead || key_masked;
              if (p.drop_thr != 0) {
                if (bitsT_row != nullptr) keep_scale = ((mw >> i) & 1u) ? p.drop_inv_keep : 0.f;
                else keep_scale = attn_drop_keep(drop_base, (uint32_t)qi, (uint32_t)k_in_seg, (uint32_t)p.Lk, p.drop_thr) ? p.drop_inv_keep : 0.f;
              }
            }
            const float pe = dead ? 0.f : ex2_approx(t - stat[cc + i]);
            pv[i] = pe * keep_scale;                                   // dropped probabilities feed dV
            dsv[i] = dead ? 0.f : pe * (dp * keep_scale - stat[BI + cc + i]);
          }
        }
        uint32_t pp[16], dd[16];
        pack16<DT>(pv, pp);
        pack16<DT>(dsv, dd);
        if (tr) TRACE(64 + j * 8 + 3);
        // P^T / dS^T were last read by the dV/dK MMAs of the previous block (the other group's)
        if (jb > 0) mbar_wait(&acc_done[g ^ 1], ((jb - 1) >> 1) & 1, 31);
        tc_fence_after();
        if (tr) TRACE(64 + j * 8 + 4);
        tmem_st_x16(tm_P + lane_off + (cc >> 1), pp);
        tmem_st_x16(tm_dS + lane_off + (cc >> 1), dd);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&pds_full[g]);
        if (tr) TRACE(64 + j * 8 + 5);
      }
      // Item transition.  All global loads go out first (next item's K / V rows, this item's inverse-RoPE rows), then the
      // barrier waits, then the staged copies: the load latency hides behind the tail of the item's MMAs.
      const int next = item + gridDim.x;
      const bool trt = threadIdx.x == 0 || threadIdx.x == 256;
      if (trt) TRACE(8 + g * 8 + 0);
      const int ch = (wi >> 2) * 32;
      const int64_t row0 = (int64_t)t_row0 + lq * 32;
      const bool do_rope = g == 1 && p.rope != nullptr;
      if (next < n_items) {
        const int ntile = next % p.tiles, nseg_ = (next / p.tiles) % p.nseg, nhead = next / (p.tiles * p.nseg);
        const int ncol = (g == 0 ? p.k_col0 : p.v_col0) + nhead * HD + ch;
        stage_fill_rows_async(stg, lane, [&](int rl) {
          const int64_t grow = (int64_t)nseg_ * p.Lk + min(ntile * BT + lq * 32 + rl, p.Lk - 1);
          return reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint16_t*>(p.kv) + grow * p.ldkv + ncol);
        });
      }
      if (do_rope) rope_fill_async(p, srope, lane, row0, ch);
      if (next < n_items) {
        mbar_wait(kv_free, it & 1, 34);
        tc_fence_after();
        if (trt) TRACE(8 + g * 8 + 1);
        stage_async_wait();
        __syncwarp();
        park_staged(stg, lane, (g == 0 ? tm_K : tm_V) + lane_off + (wi >> 2) * 16);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(kv_ready);
      }
      if (trt) TRACE(8 + g * 8 + 2);
      mbar_wait(all_done, it & 1, 32);
      tc_fence_after();
      stage_async_wait();
      __syncwarp();
      if (trt) TRACE(8 + g * 8 + 3);
      // group 0 drains dV, group 1 drains dK (+ inverse rotation): each warp its 32 rows x one 32-column half
      const int rows_valid = max(0, min(32, p.Lk - (tile * BT + lq * 32)));
      if (g == 0) store_grad_rows<DT, false>(p, stg, srope, lane, tm_dV + lane_off + ch, p.dkv, p.lddkv, row0, p.dv_col0 + head * HD + ch, 1.f, rows_valid);
      else store_grad_rows<DT, true>(p, stg, srope, lane, tm_dK + lane_off + ch, p.dkv, p.lddkv, row0, p.dk_col0 + head * HD + ch, p.scale, rows_valid);
      if (trt) TRACE(8 + g * 8 + 4);
      tc_fence_before();
      mbar_arrive_relaxed(epi_done);   // payload = drained TMEM accumulators, not the global stores above
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, TCOLS);
}

// =========================================================================================
// dQ
// =========================================================================================
// TMEM columns: S x2 [0,128)  dP x2 [128,256)  dQ [256,320)  Q [320,352)  dO [352,384)  dS x2 [384,448)
constexpr int DQ_MAIN = 2 * NSI * I_BYTES + 256;   // K ring | V ring | barriers
constexpr int DQ_SMEM = DQ_MAIN + 16 * STG_BYTES + 8 * 2 * STG_BYTES;   // + a row tile per compute warp, + 2 rope tiles per dQ warp
static_assert(DKDV_SMEM <= 227 * 1024 && DQ_SMEM <= 227 * 1024, "attention backward exceeds the 227 KB of shared memory per CTA");

// DROPB (GEN = false): the ViT instantiation + dropout through precomputed keep-bits, full blocks only (see attn_fwd.cu)
template <int DT, bool GEN, bool DROPB = false>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmI,   // kv buffer, box [64][64]
                   const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sK = smem_raw;                // NSI x 8 KB
  uint8_t* sV = sK + NSI * I_BYTES;      // NSI x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NSI * I_BYTES);
  uint64_t* in_full = bars;              // [NSI]
  uint64_t* in_free = in_full + NSI;     // [NSI]
  uint64_t* sdp_full = in_free + NSI;    // [2]
  uint64_t* sdp_free = sdp_full + 2;     // [2]
  uint64_t* ds_full = sdp_free + 2;      // [2]
  uint64_t* dq_done = ds_full + 2;       // [2] dQ MMA of a block of group b complete (dS buffer b free again)
  uint64_t* qdo_ready = dq_done + 2;
  uint64_t* qdo_free = qdo_ready + 1;
  uint64_t* all_done = qdo_free + 1;
  uint64_t* epi_done = all_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 1);

  const int warp = threadIdx.x >> 5;
  const int n_blocks = (p.Lk + BI - 1) / BI;     // 64-key blocks per item
  const int n_items = p.tiles * p.nseg * p.H;

  if (warp == WARP_TMA && elect_one()) {
    tma_prefetch_desc(&tmI);
    for (int i = 0; i < NSI; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&sdp_free[i], 256); mbar_init(&ds_full[i], 256); mbar_init(&dq_done[i], 1); }
    mbar_init(qdo_ready, 512); mbar_init(qdo_free, 1); mbar_init(all_done, 1); mbar_init(epi_done, 512);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dQ = tmem_base + 256;
  const uint32_t tm_Q = tmem_base + 320, tm_dO = tmem_base + 352, tm_dS = tmem_base + 384;

  if (warp == WARP_TMA) {
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int seg = (item / p.tiles) % p.nseg, head = item / (p.tiles * p.nseg);
        const int kv_row0_seg = seg * p.Lk;
        for (int j = 0; j < n_blocks; ++j) {
          const int jb = it * n_blocks + j, st = jb % NSI;
          if (jb >= NSI) mbar_wait(&in_free[st], (jb / NSI - 1) & 1, 10);
          mbar_arrive_expect_tx(&in_full[st], 2 * I_BYTES);
          tma_load_2d(sK + st * I_BYTES, &tmI, &in_full[st], p.k_col0 + head * HD, kv_row0_seg + j * BI);
          tma_load_2d(sV + st * I_BYTES, &tmI, &in_full[st], p.v_col0 + head * HD, kv_row0_seg + j * BI);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    if (elect_one()) {   // MMA stream 1: S, dP (see attn_bwd_dkdv_kernel for the two-stream split)
      constexpr uint32_t idesc_kk = make_idesc_f16(BT, BI, DT, 0, 0);
      constexpr uint32_t ST16 = I_BYTES >> 4;
      const uint64_t dKk = make_desc_kmajor(smem_u32(sK)), dVk = make_desc_kmajor(smem_u32(sV));
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int B0 = it * n_blocks;
        mbar_wait(qdo_ready, it & 1, 20);
        for (int j = 0; j < n_blocks; ++j) {
          const int jb = B0 + j, st = jb % NSI, b = jb & 1;
          mbar_wait(&in_full[st], (jb / NSI) & 1, 21);
          if (jb >= 2) mbar_wait(&sdp_free[b], ((jb >> 1) - 1) & 1, 24);
          tc_fence_after();
          const uint64_t kd = dKk + (uint64_t)(st * ST16), vd = dVk + (uint64_t)(st * ST16);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // S = Q . K^T
            umma_f16_ts(tm_S + b * 64, tm_Q + k * 8, kd + k * 2, idesc_kk, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dP = dO . V^T
            umma_f16_ts(tm_dP + b * 64, tm_dO + k * 8, vd + k * 2, idesc_kk, k > 0);
          umma_commit(&sdp_full[b]);
          if (j == n_blocks - 1) umma_commit(qdo_free);
        }
      }
    }
  } else if (warp == WARP_MMA2) {
    if (elect_one()) {   // MMA stream 2: dQ += dS . K
      constexpr uint32_t idesc_kmn = make_idesc_f16(BT, HD, DT, 0, 1);
      constexpr uint32_t ST16 = I_BYTES >> 4;
      const uint64_t dKm = make_desc_mnmajor(smem_u32(sK), 8192);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int B0 = it * n_blocks;
        for (int j = 0; j < n_blocks; ++j) {
          const int jb = B0 + j, st = jb % NSI, b = jb & 1;
          mbar_wait(&ds_full[b], (jb >> 1) & 1, 22);
          if (j == 0 && it > 0) mbar_wait(epi_done, (it - 1) & 1, 25);
          tc_fence_after();
          const uint64_t kd = dKm + (uint64_t)(st * ST16);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dQ += dS . K   (K = 64 keys, K tile as MN-major B)
            umma_f16_ts(tm_dQ, tm_dS + b * 32 + k * 8, kd + k * 128, idesc_kmn, (j > 0 || k > 0));
          umma_commit(&in_free[st]);
          umma_commit(&dq_done[b]);
          if (j == n_blocks - 1) umma_commit(all_done);
        }
      }
    }
  } else if (warp < 16) {
    const int g = warp >> 3;
    const int wi = warp & 7;
    const int r = threadIdx.x & 127;   // query row inside the tile == TMEM lane
    const int cc = (wi >> 2) * 32;     // this warp's half of the 64 inner (key) columns
    const uint32_t lane_off = static_cast<uint32_t>((wi & 3) * 32) << 16;
    const float c = p.scale_log2;

    uint8_t* stg = smem_raw + DQ_MAIN + warp * STG_BYTES;
    uint8_t* srope = smem_raw + DQ_MAIN + 16 * STG_BYTES + wi * 2 * STG_BYTES;     // used by group 0 (dQ) only
    const int lane = threadIdx.x & 31, lq = wi & 3;

    // Q (group 0) / dO (group 1) rows of an item -> TMEM (coalesced through the warp's staging tile)
    auto park = [&](int item) {
      const int tile = item % p.tiles, seg = (item / p.tiles) % p.nseg, head = item / (p.tiles * p.nseg);
      const uint16_t* base = reinterpret_cast<const uint16_t*>(g == 0 ? p.q : p.dO);
      const int64_t ldb = g == 0 ? p.ldq : p.lddo;
      const int col = (g == 0 ? p.q_col0 : p.do_col0) + head * HD + (wi >> 2) * 32;
      park_rows(stg, lane, (g == 0 ? tm_Q : tm_dO) + lane_off + (wi >> 2) * 16, [&](int rl) {
        const int64_t grow = (int64_t)seg * p.Lq + min(tile * BT + lq * 32 + rl, p.Lq - 1);
        return reinterpret_cast<const uint8_t*>(base + grow * ldb + col);
      });
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(qdo_ready);
    };
    if ((int)blockIdx.x < n_items) park(blockIdx.x);

    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item % p.tiles, seg = (item / p.tiles) % p.nseg, head = item / (p.tiles * p.nseg);
      const int B0 = it * n_blocks;
      const int t_row0 = seg * p.Lq + tile * BT;   // this item's 128 queries
      const int q_in_seg = min(tile * BT + r, p.Lq - 1);
      const int64_t sidx = (int64_t)head * p.stat_stride + (int64_t)seg * p.Lq_stat + q_in_seg;   // head-major statistics
      const float lse = __ldg(p.lse2 + sidx);
      const float dlt = __ldg(p.delta + sidx);
      const float* bias_row = nullptr;
      const uint8_t* kpm_row = nullptr;
      uint32_t drop_base = 0;
      if constexpr (GEN) {
        if (p.bias != nullptr) bias_row = p.bias + ((int64_t)(seg * p.H + head) * p.Lq + q_in_seg) * p.Lk;
        if (p.kpm != nullptr) kpm_row = p.kpm + (int64_t)seg * p.Lk;
        drop_base = attn_drop_base(p.drop_seed, (uint32_t)(seg * p.H + head));
      }
      const uint32_t* bits_row = nullptr;
      if constexpr (GEN || DROPB) {
        if (p.drop_bits != nullptr) bits_row = p.drop_bits + ((int64_t)(seg * p.H + head) * p.Lq + q_in_seg) * p.bits_pitch_k;
      }
      for (int j = (B0 + g) & 1; j < n_blocks; j += 2) {
        const int jb = B0 + j;
        const int k_valid = min(BI, p.Lk - j * BI);
        mbar_wait(&sdp_full[g], (jb >> 1) & 1, 30);
        tc_fence_after();
        uint32_t s[32], d[32];
        tmem_ld_x32(tm_S + g * 64 + lane_off + cc, s);
        tmem_ld_x32(tm_dP + g * 64 + lane_off + cc, d);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sdp_free[g]);
        float dsv[32];
        bool fast = k_valid == BI;
        if constexpr (GEN) fast = fast && bias_row == nullptr && kpm_row == nullptr && (p.drop_thr == 0 || bits_row != nullptr);
        if (fast) {
          uint32_t mwf = 0xffffffffu;
          float inv_keep = 1.f;
          if constexpr (GEN || DROPB) {
            if (DROPB || p.drop_thr != 0) { mwf = __ldg(bits_row + ((j * BI + cc) >> 5)); inv_keep = p.drop_inv_keep; }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pe = ex2_approx(fmaf(__uint_as_float(s[i]), c, -lse));
            if constexpr (GEN || DROPB) dsv[i] = pe * fmaf(__uint_as_float(d[i]), ((mwf >> i) & 1u) ? inv_keep : 0.f, -dlt);
            else dsv[i] = pe * (__uint_as_float(d[i]) - dlt);
          }
        } else {
          uint32_t mw = 0u;
          bool bias_pre = false;        // the row's 32 bias values come in through 16-byte loads (see attn_fwd.cu)
          if constexpr (GEN) {
            if (bits_row != nullptr) { const int wi_ = (j * BI + cc) >> 5; mw = wi_ < p.bits_pitch_k ? __ldg(bits_row + wi_) : 0u; }
            bias_pre = bias_row != nullptr && p.bias_vec4 && cc + 32 <= k_valid;
          }
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            float4 b4v = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (GEN) {
              if (bias_pre) b4v = __ldg(reinterpret_cast<const float4*>(bias_row + j * BI + cc) + q4);
            }
            const float bq[4] = {b4v.x, b4v.y, b4v.z, b4v.w};
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int i = q4 * 4 + e4;
              const int ki = j * BI + cc + i;   // key index inside the segment
              float t = __uint_as_float(s[i]) * c;
              float dp = __uint_as_float(d[i]);
              bool dead = (cc + i) >= k_valid;
              if constexpr (GEN) {
                if (bias_pre) t = fmaf(bq[e4], 1.4426950408889634f, t);
                else if (bias_row != nullptr && !dead) t = fmaf(__ldg(bias_row + ki), 1.4426950408889634f, t);
                if (kpm_row != nullptr && !dead) dead = __ldg(kpm_row + ki) != 0;
                if (p.drop_thr != 0) {
                  if (bits_row != nullptr) dp *= ((mw >> i) & 1u) ? p.drop_inv_keep : 0.f;
                  else dp *= attn_drop_keep(drop_base, (uint32_t)q_in_seg, (uint32_t)ki, (uint32_t)p.Lk, p.drop_thr) ? p.drop_inv_keep : 0.f;
                }
              }
              const float pe = dead ? 0.f : ex2_approx(t - lse);
              dsv[i] = pe * (dp - dlt);
            }
          }
        }
        uint32_t dd[16];
        pack16<DT>(dsv, dd);
        // dS buffer g was last read by the dQ MMA of this group's previous block
        if (jb >= 2) mbar_wait(&dq_done[g], ((jb >> 1) - 1) & 1, 31);
        tc_fence_after();
        tmem_st_x16(tm_dS + g * 32 + lane_off + (cc >> 1), dd);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ds_full[g]);
      }
      // item transition: loads first (next item's Q / dO rows, this item's inverse-RoPE rows), then waits, then copies
      const int next = item + gridDim.x;
      const int ch = (wi >> 2) * 32;
      const int64_t row0 = (int64_t)t_row0 + lq * 32;
      const bool do_rope = g == 0 && p.rope != nullptr;
      if (next < n_items) {
        const int ntile = next % p.tiles, nseg_ = (next / p.tiles) % p.nseg, nhead = next / (p.tiles * p.nseg);
        const uint16_t* base = reinterpret_cast<const uint16_t*>(g == 0 ? p.q : p.dO);
        const int64_t ldb = g == 0 ? p.ldq : p.lddo;
        const int ncol = (g == 0 ? p.q_col0 : p.do_col0) + nhead * HD + ch;
        stage_fill_rows_async(stg, lane, [&](int rl) {
          const int64_t grow = (int64_t)nseg_ * p.Lq + min(ntile * BT + lq * 32 + rl, p.Lq - 1);
          return reinterpret_cast<const uint8_t*>(base + grow * ldb + ncol);
        });
      }
      if (do_rope) rope_fill_async(p, srope, lane, row0, ch);
      if (next < n_items) {
        mbar_wait(qdo_free, it & 1, 34);
        tc_fence_after();
        stage_async_wait();
        __syncwarp();
        park_staged(stg, lane, (g == 0 ? tm_Q : tm_dO) + lane_off + (wi >> 2) * 16);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(qdo_ready);
      }
      mbar_wait(all_done, it & 1, 32);
      tc_fence_after();
      stage_async_wait();
      __syncwarp();
      if (g == 0) {   // 8 warps x (32 rows x 32 columns): whole 64-byte row segments per lane group
        const int rows_valid = max(0, min(32, p.Lq - (tile * BT + lq * 32)));
        store_grad_rows<DT, true>(p, stg, srope, lane, tm_dQ + lane_off + ch, p.dq, p.lddq, row0, p.dq_col0 + head * HD + ch, p.scale, rows_valid);
      }
      tc_fence_before();
      mbar_arrive_relaxed(epi_done);   // payload = drained TMEM accumulators, not the global stores above
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, TCOLS);
}

template <typename K>
static int set_smem_once(K kern, int bytes, bool& done) {
  if (!done) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done = true;
  }
  return 0;
}

template <int DT, bool GEN, bool DROPB = false>
static int launch_bwd(const AttnArgs& a, const BwdParams& pk, const BwdParams& pq, const CUtensorMap& tmQ64, const CUtensorMap& tmdO64,
                      const CUtensorMap& tmKV64, cudaStream_t stream) {
  static bool s1 = false, s2 = false;
  int rc = set_smem_once(attn_bwd_dkdv_kernel<DT, GEN, DROPB>, DKDV_SMEM, s1);
  if (rc) return rc;
  rc = set_smem_once(attn_bwd_dq_kernel<DT, GEN, DROPB>, DQ_SMEM, s2);
  if (rc) return rc;
  // persistent: one CTA per SM walks the item list (SAM3B_ATTN_PERSIST=0: one CTA per item, for debugging)
  static const bool persist = [] { const char* e = getenv("SAM3B_ATTN_PERSIST"); return !(e && e[0] == '0'); }();
  const int items_k = pk.tiles * a.nseg * a.heads, items_q = pq.tiles * a.nseg * a.heads;
  const int grid_k = persist ? std::min(items_k, num_sms()) : items_k;
  const int grid_q = persist ? std::min(items_q, num_sms()) : items_q;
  SAM3B_CHECK_CUDA(launch_pdl(attn_bwd_dkdv_kernel<DT, GEN, DROPB>, dim3(grid_k), dim3(NTHREADS), DKDV_SMEM, stream, tmQ64, tmdO64, pk));
  SAM3B_LAUNCHED();
  SAM3B_CHECK_CUDA(launch_pdl(attn_bwd_dq_kernel<DT, GEN, DROPB>, dim3(grid_q), dim3(NTHREADS), DQ_SMEM, stream, tmKV64, pq));
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace

int attn_bwd_launch(const AttnArgs& a, cudaStream_t stream) {
  SAM3B_REQUIRE(a.q && a.kv && a.dO && a.lse2 && a.delta && a.dq && a.dkv, "attention bwd: null tensor");
  SAM3B_REQUIRE(a.nseg > 0 && a.Lq > 0 && a.Lk > 0 && a.heads > 0, "attention bwd: empty problem");
  SAM3B_REQUIRE(a.ldq % 8 == 0 && a.ldkv % 8 == 0 && a.lddo % 8 == 0 && a.lddq % 8 == 0 && a.lddkv % 8 == 0,
                "attention bwd: leading dimensions must be multiples of 8");
  SAM3B_REQUIRE(a.drop_p >= 0.f && a.drop_p < 1.f, "attention bwd: dropout p");
  const uint64_t Mq = (uint64_t)a.nseg * a.Lq, Mk = (uint64_t)a.nseg * a.Lk;
  SAM3B_REQUIRE(a.q_col0 % 8 == 0 && a.k_col0 % 8 == 0 && a.v_col0 % 8 == 0 && a.do_col0 % 8 == 0,
                "attention bwd: column offsets must be multiples of 8");
  CUtensorMap tmQ64, tmdO64, tmKV64;
  int rc;
  if ((rc = make_tmap_2d(&tmKV64, a.kv, Mk, a.kv_cols, a.ldkv, BI, HD))) return rc;
  if ((rc = make_tmap_2d(&tmQ64, a.q, Mq, a.q_cols, a.ldq, BI, HD))) return rc;
  const int do_cols = a.do_col0 + a.heads * HD;
  if ((rc = make_tmap_2d(&tmdO64, a.dO, Mq, do_cols, a.lddo, BI, HD))) return rc;
  BwdParams p{};
  p.Lq = a.Lq; p.Lk = a.Lk; p.H = a.heads; p.nseg = a.nseg;
  p.q = a.q; p.ldq = a.ldq; p.kv = a.kv; p.ldkv = a.ldkv; p.dO = a.dO; p.lddo = a.lddo;
  p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0; p.do_col0 = a.do_col0;
  p.lse2 = a.lse2; p.delta = a.delta; p.Lq_stat = attn_lq_stat(a.Lq); p.stat_stride = (int64_t)a.nseg * p.Lq_stat;
  p.dq = a.dq; p.lddq = a.lddq; p.dq_col0 = a.dq_col0;
  p.dkv = a.dkv; p.lddkv = a.lddkv; p.dk_col0 = a.dk_col0; p.dv_col0 = a.dv_col0;
  p.rope = reinterpret_cast<const float2*>(a.rope); p.rope_period = a.rope_period > 0 ? a.rope_period : 1;
  p.scale = a.scale; p.scale_log2 = a.scale * 1.4426950408889634f;
  p.bias = a.bias; p.kpm = a.kpm;
  p.drop_inv_keep = 1.f / (1.f - a.drop_p); p.drop_thr = a.drop_p > 0.f ? dropout_threshold(a.drop_p) : 0u; p.drop_seed = a.drop_seed;
  {
    const bool bits_ok = a.drop_p > 0.f && a.Lq % 32 == 0 && a.Lk % 32 == 0 && a.drop_bits != nullptr && a.drop_bitsT != nullptr;
    p.drop_bits = bits_ok ? a.drop_bits : nullptr;
    p.drop_bitsT = bits_ok ? a.drop_bitsT : nullptr;
  }
  p.bits_pitch_k = attn_bits_pitch(a.Lk); p.bits_pitch_q = attn_bits_pitch(a.Lq);
  p.bias_vec4 = (a.bias != nullptr && a.Lk % 4 == 0 && (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) ? 1 : 0;
  BwdParams pk = p, pq = p;
  pk.tiles = (a.Lk + BT - 1) / BT;
  pq.tiles = (a.Lq + BT - 1) / BT;
  const bool gen = a.bias != nullptr || a.kpm != nullptr || a.drop_p > 0.f;
  if (a.bias == nullptr && a.kpm == nullptr && a.drop_p > 0.f && p.drop_bits != nullptr && a.Lq == a.Lk && a.Lk % BI == 0)
    return a.dtype == 0 ? launch_bwd<0, false, true>(a, pk, pq, tmQ64, tmdO64, tmKV64, stream)
                        : launch_bwd<1, false, true>(a, pk, pq, tmQ64, tmdO64, tmKV64, stream);
  if (gen)
    return a.dtype == 0 ? launch_bwd<0, true>(a, pk, pq, tmQ64, tmdO64, tmKV64, stream)
                        : launch_bwd<1, true>(a, pk, pq, tmQ64, tmdO64, tmKV64, stream);
  return a.dtype == 0 ? launch_bwd<0, false>(a, pk, pq, tmQ64, tmdO64, tmKV64, stream)
                      : launch_bwd<1, false>(a, pk, pq, tmQ64, tmdO64, tmKV64, stream);
}

#ifdef SAM3B_TRACE
int attn_trace_read(unsigned long long* host, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(unsigned long long) * (size_t)n);
}
int attn_trace_clear() {
  static unsigned long long z[16384];
  return (int)cudaMemcpyToSymbol(g_attn_trace, z, sizeof(z));
}
#endif

}  // namespace sam3b
