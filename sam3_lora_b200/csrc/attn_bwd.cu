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
//   attn_bwd_dkdv : one CTA per (128-key tile, head, segment), loops over 64-query blocks.
//                   S^T = K Q^T and dP^T = V dO^T in TMEM (thread = key row), P^T / dS^T written
//                   16-bit to swizzled smem as A operands, dV += P^T dO, dK += dS^T Q with the
//                   same Q/dO tiles re-used as MN-major B operands.
//   attn_bwd_dq   : one CTA per (128-query tile, head, segment), loops over 64-key blocks.
//                   S = Q K^T, dP = dO V^T (thread = query row), dQ += dS K (K tile as MN-major B).
// Both: warps 0-7 compute (two warps per TMEM lane quarter, each owning 32 of the 64 inner columns, so
// 16 compute warps per SM hide the tcgen05.ld / MUFU latencies), warp 8 TMA, warp 9 MMA issue + TMEM
// alloc; 256 TMEM columns and <= 100 KB smem so two CTAs share an SM.
#include "attn.cuh"

#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace sam3b {

#ifdef SAM3B_TRACE
// Debug timeline (tools/attn_trace.py): one chosen CTA stamps clock64() at its pipeline events.
__device__ unsigned long long g_attn_trace[16384];
#define TR_CTA 1500
#define TRACE(slot) do { if (blockIdx.x + blockIdx.y * gridDim.x == TR_CTA && (slot) < 16384) g_attn_trace[(slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

namespace {

constexpr int HD = 64;
constexpr int BT = 128;  // rows owned by a CTA (keys in dkdv, queries in dq)
constexpr int BI = 64;   // inner block (queries in dkdv, keys in dq)
constexpr int T_BYTES = BT * HD * 2;   // 16 KB
constexpr int I_BYTES = BI * HD * 2;   // 8 KB
constexpr int A_BYTES = BT * BI * 2;   // 16 KB: [128][64] 16-bit A operand written by the compute warps
constexpr int TCOLS = 256;

struct BwdParams {
  int Lq, Lk, tiles, H;
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
  float drop_inv_keep; uint32_t drop_thr, drop_seed;
};

__device__ __forceinline__ uint32_t attn_drop_base(uint32_t seed, uint32_t bh) { return lowbias32(seed ^ (bh * 0x9E3779B1u + 0x85EBCA6Bu)); }
__device__ __forceinline__ bool attn_drop_keep(uint32_t base, uint32_t q, uint32_t k, uint32_t Lk, uint32_t thr) {
  return lowbias32(base ^ (q * Lk + k)) >= thr;
}

template <int DT>
__device__ __forceinline__ void store_row_chunk16(uint8_t* row_base, int sw, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack2<DT>(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack2<DT>(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack2<DT>(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack2<DT>(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(row_base + (((ch0 + q) ^ sw) << 4)) = u;
  }
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

// =========================================================================================
// dK / dV
// =========================================================================================
constexpr int DKDV_SMEM = 2 * T_BYTES + 2 * (2 * I_BYTES) + 2 * A_BYTES + 2 * 2 * BI * 4 + 128;  // stats: [2 stages][lse2 | delta][64]

template <int DT, bool GEN>
__global__ void __launch_bounds__(320, 2)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmT,   // kv buffer, box [128][64]
                     const __grid_constant__ CUtensorMap tmI,   // q buffer,  box [64][64]
                     const __grid_constant__ CUtensorMap tmdO,  // dO,        box [64][64]
                     const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sK = smem_raw;
  uint8_t* sV = sK + T_BYTES;
  uint8_t* sQ = sV + T_BYTES;            // 2 stages x 8 KB
  uint8_t* sdO = sQ + 2 * I_BYTES;       // 2 stages x 8 KB
  uint8_t* sP = sdO + 2 * I_BYTES;       // P^T  [128 keys][64 q]
  uint8_t* sdS = sP + A_BYTES;           // dS^T [128 keys][64 q]
  float* sStat = reinterpret_cast<float*>(sdS + A_BYTES);  // [2 stages][2][64]: lse2 | delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 2 * BI);
  uint64_t* kv_full = bars + 0;
  uint64_t* in_full = bars + 1;   // [2]
  uint64_t* in_free = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* acc_done = bars + 7;  // dV/dK MMAs of block j complete (also frees sP/sdS)
  uint64_t* sdp_free = bars + 8;  // compute warps have S^T/dP^T of block j in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) TRACE(0);
  const int tile = blockIdx.x % p.tiles;
  const int seg = blockIdx.x / p.tiles;
  const int head = blockIdx.y;
  const int q_row0_seg = seg * p.Lq;
  const int t_row0 = seg * p.Lk + tile * BT;   // this CTA's 128 keys
  const int n_blocks = (p.Lq + BI - 1) / BI;     // 64-query blocks

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmI); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_free[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, 256); mbar_init(acc_done, 1); mbar_init(sdp_free, 256);
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TRACE(1);
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 64, tm_dV = tmem_base + 128, tm_dK = tmem_base + 192;

  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * T_BYTES);
      tma_load_2d(sK, &tmT, kv_full, p.k_col0 + head * HD, t_row0);
      tma_load_2d(sV, &tmT, kv_full, p.v_col0 + head * HD, t_row0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&in_free[st], ((j >> 1) - 1) & 1, 10 + st);
        mbar_arrive_expect_tx(&in_full[st], 2 * I_BYTES + 2 * BI * 4);
        tma_load_2d(sQ + st * I_BYTES, &tmI, &in_full[st], p.q_col0 + head * HD, q_row0_seg + j * BI);
        tma_load_2d(sdO + st * I_BYTES, &tmdO, &in_full[st], p.do_col0 + head * HD, q_row0_seg + j * BI);
        // per-query statistics of this block (head-major layout: 64 consecutive floats each)
        const int64_t soff = (int64_t)head * p.stat_stride + (int64_t)seg * p.Lq_stat + j * BI;
        bulk_load_1d(sStat + st * (2 * BI), p.lse2 + soff, BI * 4, &in_full[st]);
        bulk_load_1d(sStat + st * (2 * BI) + BI, p.delta + soff, BI * 4, &in_full[st]);
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_f16(BT, BI, DT, 0, 0);  // A K-major, B K-major, N=64
      constexpr uint32_t idesc_kmn = make_idesc_f16(BT, HD, DT, 0, 1); // A K-major, B MN-major, N=64
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      mbar_wait(kv_full, 0, 20);
      TRACE(2);
      // S^T / dP^T of block j+1 are issued as soon as the compute warps hold block j in registers
      // (sdp_free), i.e. they overlap the exp / pack / store work of block j instead of waiting for it.
      auto issue_sdp = [&](int j) {
        const int st = j & 1;
        const uint32_t q_addr = smem_u32(sQ + st * I_BYTES), do_addr = smem_u32(sdO + st * I_BYTES);
        mbar_wait(&in_full[st], (j >> 1) & 1, 21);
        if (j > 0) mbar_wait(sdp_free, (j - 1) & 1, 24);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T = K . Q^T
          umma_f16_ss(tm_S, make_desc_kmajor(k_addr + k * 32), make_desc_kmajor(q_addr + k * 32), idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T = V . dO^T
          umma_f16_ss(tm_dP, make_desc_kmajor(v_addr + k * 32), make_desc_kmajor(do_addr + k * 32), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      issue_sdp(0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        const uint32_t q_addr = smem_u32(sQ + st * I_BYTES), do_addr = smem_u32(sdO + st * I_BYTES);
        TRACE(64 + j * 16 + 8);
        if (j + 1 < n_blocks) issue_sdp(j + 1);
        TRACE(64 + j * 16 + 9);
        mbar_wait(pds_full, j & 1, 22);
        TRACE(64 + j * 16 + 10);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dV += P^T . dO   (K = 64 queries)
          umma_f16_ss(tm_dV, make_desc_kmajor(p_addr + k * 32), make_desc_mnmajor(do_addr + k * 2048, 8192), idesc_kmn,
                      (j > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dK += dS^T . Q
          umma_f16_ss(tm_dK, make_desc_kmajor(ds_addr + k * 32), make_desc_mnmajor(q_addr + k * 2048, 8192), idesc_kmn,
                      (j > 0 || k > 0));
        umma_commit(&in_free[st]);
        umma_commit(acc_done);
        TRACE(64 + j * 16 + 11);
      }
    }
  } else {
    const int r = threadIdx.x & 127;   // key row inside the tile == TMEM lane
    const int cc = (warp >> 2) * 32;   // this warp's half of the 64 inner (query) columns
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float c = p.scale_log2;
    uint8_t* p_row = sP + r * 128;
    uint8_t* ds_row = sdS + r * 128;
    const int sw = r & 7;
    const int k_in_seg = min(tile * BT + r, p.Lk - 1);
    bool key_masked = false;
    const float* bias_col = nullptr;
    uint32_t drop_base = 0;
    if constexpr (GEN) {
      if (p.kpm != nullptr) key_masked = __ldg(p.kpm + (int64_t)seg * p.Lk + k_in_seg) != 0;
      if (p.bias != nullptr) bias_col = p.bias + (int64_t)(seg * p.H + head) * p.Lq * p.Lk + k_in_seg;
      drop_base = attn_drop_base(p.drop_seed, (uint32_t)(seg * p.H + head));
    }
    for (int j = 0; j < n_blocks; ++j) {
      // per-query statistics were bulk-copied next to Q/dO by the producer (visible once in_full completed)
      const float* stat = sStat + (j & 1) * (2 * BI);
      const int q_valid = min(BI, p.Lq - j * BI);   // queries of this block that exist
      if (threadIdx.x == 0) TRACE(64 + j * 16 + 0);
      mbar_wait(&in_full[j & 1], (j >> 1) & 1, 33);
      mbar_wait(sdp_full, j & 1, 30);
      if (threadIdx.x == 0) TRACE(64 + j * 16 + 1);
      tc_fence_after();
      {
        uint32_t s[32], d[32];
        tmem_ld_x32(tm_S + lane_off + cc, s);
        tmem_ld_x32(tm_dP + lane_off + cc, d);
        tmem_ld_wait();
        if (threadIdx.x == 0) TRACE(64 + j * 16 + 2);
        tc_fence_before();
        mbar_arrive(sdp_free);  // TMEM S^T/dP^T may be overwritten by block j+1
        float pv[32], dsv[32];
        if (!GEN && q_valid == BI) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pe = ex2_approx(__uint_as_float(s[i]) * c - stat[cc + i]);
            pv[i] = pe;
            dsv[i] = pe * (__uint_as_float(d[i]) - stat[BI + cc + i]);
          }
        } else {
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
              if (p.drop_thr != 0)
                keep_scale = attn_drop_keep(drop_base, (uint32_t)qi, (uint32_t)k_in_seg, (uint32_t)p.Lk, p.drop_thr) ? p.drop_inv_keep : 0.f;
            }
            const float pe = dead ? 0.f : ex2_approx(t - stat[cc + i]);
            pv[i] = pe * keep_scale;                                   // dropped probabilities feed dV
            dsv[i] = dead ? 0.f : pe * (dp * keep_scale - stat[BI + cc + i]);
          }
        }
        if (threadIdx.x == 0) TRACE(64 + j * 16 + 3);
        if (j > 0) mbar_wait(acc_done, (j - 1) & 1, 31);
        if (threadIdx.x == 0) TRACE(64 + j * 16 + 4);  // previous dV/dK MMAs finished reading sP/sdS
        store_row_chunk16<DT>(p_row, sw, cc >> 3, pv);
        store_row_chunk16<DT>(ds_row, sw, cc >> 3, dsv);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(pds_full);
      if (threadIdx.x == 0) TRACE(64 + j * 16 + 5);
    }
    mbar_wait(acc_done, (n_blocks - 1) & 1, 32);
    if (threadIdx.x == 0) TRACE(3);
    tc_fence_after();
    const int row = t_row0 + r;
    const bool valid = (tile * BT + r) < p.Lk;
    store_grad_chunk<DT, false>(p, p.dkv, p.lddkv, tm_dV + lane_off, row, row, p.dv_col0 + head * HD, cc, 1.f, valid);
    store_grad_chunk<DT, true>(p, p.dkv, p.lddkv, tm_dK + lane_off, row, row, p.dk_col0 + head * HD, cc, p.scale, valid);
    if (threadIdx.x == 0) TRACE(4);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, TCOLS);
}

// =========================================================================================
// dQ
// =========================================================================================
constexpr int DQ_SMEM = 2 * T_BYTES + 2 * (2 * I_BYTES) + A_BYTES + 128;

template <int DT, bool GEN>
__global__ void __launch_bounds__(320, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmT,   // q buffer,  box [128][64]
                   const __grid_constant__ CUtensorMap tmI,   // kv buffer, box [64][64]
                   const __grid_constant__ CUtensorMap tmdO,  // dO,        box [128][64]
                   const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sQ = smem_raw;
  uint8_t* sdO = sQ + T_BYTES;
  uint8_t* sK = sdO + T_BYTES;           // 2 stages x 8 KB
  uint8_t* sV = sK + 2 * I_BYTES;        // 2 stages x 8 KB
  uint8_t* sdS = sV + 2 * I_BYTES;       // dS [128 q][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + A_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* in_full = bars + 1;   // [2]
  uint64_t* in_free = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint64_t* sdp_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int tile = blockIdx.x % p.tiles;
  const int seg = blockIdx.x / p.tiles;
  const int head = blockIdx.y;
  const int kv_row0_seg = seg * p.Lk;
  const int t_row0 = seg * p.Lq + tile * BT;   // this CTA's 128 queries
  const int n_blocks = (p.Lk + BI - 1) / BI;     // 64-key blocks

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmI); tma_prefetch_desc(&tmdO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_free[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(ds_full, 256); mbar_init(acc_done, 1); mbar_init(sdp_free, 256);
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 64, tm_dQ = tmem_base + 128;

  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * T_BYTES);
      tma_load_2d(sQ, &tmT, q_full, p.q_col0 + head * HD, t_row0);
      tma_load_2d(sdO, &tmdO, q_full, p.do_col0 + head * HD, t_row0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&in_free[st], ((j >> 1) - 1) & 1, 10 + st);
        mbar_arrive_expect_tx(&in_full[st], 2 * I_BYTES);
        tma_load_2d(sK + st * I_BYTES, &tmI, &in_full[st], p.k_col0 + head * HD, kv_row0_seg + j * BI);
        tma_load_2d(sV + st * I_BYTES, &tmI, &in_full[st], p.v_col0 + head * HD, kv_row0_seg + j * BI);
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_f16(BT, BI, DT, 0, 0);
      constexpr uint32_t idesc_kmn = make_idesc_f16(BT, HD, DT, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS);
      mbar_wait(q_full, 0, 20);
      auto issue_sdp = [&](int j) {
        const int st = j & 1;
        const uint32_t k_addr = smem_u32(sK + st * I_BYTES), v_addr = smem_u32(sV + st * I_BYTES);
        mbar_wait(&in_full[st], (j >> 1) & 1, 21);
        if (j > 0) mbar_wait(sdp_free, (j - 1) & 1, 24);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S = Q . K^T
          umma_f16_ss(tm_S, make_desc_kmajor(q_addr + k * 32), make_desc_kmajor(k_addr + k * 32), idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP = dO . V^T
          umma_f16_ss(tm_dP, make_desc_kmajor(do_addr + k * 32), make_desc_kmajor(v_addr + k * 32), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      issue_sdp(0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        const uint32_t k_addr = smem_u32(sK + st * I_BYTES);
        if (j + 1 < n_blocks) issue_sdp(j + 1);
        mbar_wait(ds_full, j & 1, 22);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dQ += dS . K   (K = 64 keys, K tile as MN-major B)
          umma_f16_ss(tm_dQ, make_desc_kmajor(ds_addr + k * 32), make_desc_mnmajor(k_addr + k * 2048, 8192), idesc_kmn,
                      (j > 0 || k > 0));
        umma_commit(&in_free[st]);
        umma_commit(acc_done);
      }
    }
  } else {
    const int r = threadIdx.x & 127;   // query row inside the tile == TMEM lane
    const int cc = (warp >> 2) * 32;   // this warp's half of the 64 inner (key) columns
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float c = p.scale_log2;
    const int row = t_row0 + r;
    const int q_in_seg = min(tile * BT + r, p.Lq - 1);
    const int64_t sidx = (int64_t)head * p.stat_stride + (int64_t)seg * p.Lq_stat + q_in_seg;   // head-major statistics
    const float lse = __ldg(p.lse2 + sidx);
    const float dlt = __ldg(p.delta + sidx);
    uint8_t* ds_row = sdS + r * 128;
    const int sw = r & 7;
    const float* bias_row = nullptr;
    const uint8_t* kpm_row = nullptr;
    uint32_t drop_base = 0;
    if constexpr (GEN) {
      if (p.bias != nullptr) bias_row = p.bias + ((int64_t)(seg * p.H + head) * p.Lq + q_in_seg) * p.Lk;
      if (p.kpm != nullptr) kpm_row = p.kpm + (int64_t)seg * p.Lk;
      drop_base = attn_drop_base(p.drop_seed, (uint32_t)(seg * p.H + head));
    }
    for (int j = 0; j < n_blocks; ++j) {
      const int k_valid = min(BI, p.Lk - j * BI);
      mbar_wait(sdp_full, j & 1, 30);
      tc_fence_after();
      {
        uint32_t s[32], d[32];
        tmem_ld_x32(tm_S + lane_off + cc, s);
        tmem_ld_x32(tm_dP + lane_off + cc, d);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(sdp_free);
        float dsv[32];
        if (!GEN && k_valid == BI) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pe = ex2_approx(__uint_as_float(s[i]) * c - lse);
            dsv[i] = pe * (__uint_as_float(d[i]) - dlt);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ki = j * BI + cc + i;   // key index inside the segment
            float t = __uint_as_float(s[i]) * c;
            float dp = __uint_as_float(d[i]);
            bool dead = (cc + i) >= k_valid;
            if constexpr (GEN) {
              if (bias_row != nullptr && !dead) t = fmaf(__ldg(bias_row + ki), 1.4426950408889634f, t);
              if (kpm_row != nullptr && !dead) dead = __ldg(kpm_row + ki) != 0;
              if (p.drop_thr != 0)
                dp *= attn_drop_keep(drop_base, (uint32_t)q_in_seg, (uint32_t)ki, (uint32_t)p.Lk, p.drop_thr) ? p.drop_inv_keep : 0.f;
            }
            const float pe = dead ? 0.f : ex2_approx(t - lse);
            dsv[i] = pe * (dp - dlt);
          }
        }
        if (j > 0) mbar_wait(acc_done, (j - 1) & 1, 31);
        store_row_chunk16<DT>(ds_row, sw, cc >> 3, dsv);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_done, (n_blocks - 1) & 1, 32);
    tc_fence_after();
    const bool valid = (tile * BT + r) < p.Lq;
    store_grad_chunk<DT, true>(p, p.dq, p.lddq, tm_dQ + lane_off, row, row, p.dq_col0 + head * HD, cc, p.scale, valid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, TCOLS);
}

template <typename K>
static int set_smem_once(K kern, int bytes, bool& done) {
  if (!done) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done = true;
  }
  return 0;
}

template <int DT, bool GEN>
static int launch_bwd(const AttnArgs& a, const BwdParams& pk, const BwdParams& pq, const CUtensorMap& tmKV128, const CUtensorMap& tmQ64,
                      const CUtensorMap& tmdO64, const CUtensorMap& tmQ128, const CUtensorMap& tmKV64, const CUtensorMap& tmdO128,
                      cudaStream_t stream) {
  static bool s1 = false, s2 = false;
  int rc = set_smem_once(attn_bwd_dkdv_kernel<DT, GEN>, DKDV_SMEM, s1);
  if (rc) return rc;
  rc = set_smem_once(attn_bwd_dq_kernel<DT, GEN>, DQ_SMEM, s2);
  if (rc) return rc;
  attn_bwd_dkdv_kernel<DT, GEN><<<dim3(pk.tiles * a.nseg, a.heads), 320, DKDV_SMEM, stream>>>(tmKV128, tmQ64, tmdO64, pk);
  SAM3B_LAUNCHED();
  attn_bwd_dq_kernel<DT, GEN><<<dim3(pq.tiles * a.nseg, a.heads), 320, DQ_SMEM, stream>>>(tmQ128, tmKV64, tmdO128, pq);
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
  CUtensorMap tmKV128, tmQ64, tmdO64, tmQ128, tmKV64, tmdO128;
  int rc;
  if ((rc = make_tmap_2d(&tmKV128, a.kv, Mk, a.kv_cols, a.ldkv, BT, HD))) return rc;
  if ((rc = make_tmap_2d(&tmKV64, a.kv, Mk, a.kv_cols, a.ldkv, BI, HD))) return rc;
  if ((rc = make_tmap_2d(&tmQ128, a.q, Mq, a.q_cols, a.ldq, BT, HD))) return rc;
  if ((rc = make_tmap_2d(&tmQ64, a.q, Mq, a.q_cols, a.ldq, BI, HD))) return rc;
  const int do_cols = a.do_col0 + a.heads * HD;
  if ((rc = make_tmap_2d(&tmdO64, a.dO, Mq, do_cols, a.lddo, BI, HD))) return rc;
  if ((rc = make_tmap_2d(&tmdO128, a.dO, Mq, do_cols, a.lddo, BT, HD))) return rc;
  BwdParams p{};
  p.Lq = a.Lq; p.Lk = a.Lk; p.H = a.heads;
  p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0; p.do_col0 = a.do_col0;
  p.lse2 = a.lse2; p.delta = a.delta; p.Lq_stat = attn_lq_stat(a.Lq); p.stat_stride = (int64_t)a.nseg * p.Lq_stat;
  p.dq = a.dq; p.lddq = a.lddq; p.dq_col0 = a.dq_col0;
  p.dkv = a.dkv; p.lddkv = a.lddkv; p.dk_col0 = a.dk_col0; p.dv_col0 = a.dv_col0;
  p.rope = reinterpret_cast<const float2*>(a.rope); p.rope_period = a.rope_period > 0 ? a.rope_period : 1;
  p.scale = a.scale; p.scale_log2 = a.scale * 1.4426950408889634f;
  p.bias = a.bias; p.kpm = a.kpm;
  p.drop_inv_keep = 1.f / (1.f - a.drop_p); p.drop_thr = a.drop_p > 0.f ? dropout_threshold(a.drop_p) : 0u; p.drop_seed = a.drop_seed;
  BwdParams pk = p, pq = p;
  pk.tiles = (a.Lk + BT - 1) / BT;
  pq.tiles = (a.Lq + BT - 1) / BT;
  const bool gen = a.bias != nullptr || a.kpm != nullptr || a.drop_p > 0.f;
  if (gen)
    return a.dtype == 0 ? launch_bwd<0, true>(a, pk, pq, tmKV128, tmQ64, tmdO64, tmQ128, tmKV64, tmdO128, stream)
                        : launch_bwd<1, true>(a, pk, pq, tmKV128, tmQ64, tmdO64, tmQ128, tmKV64, tmdO128, stream);
  return a.dtype == 0 ? launch_bwd<0, false>(a, pk, pq, tmKV128, tmQ64, tmdO64, tmQ128, tmKV64, tmdO128, stream)
                      : launch_bwd<1, false>(a, pk, pq, tmKV128, tmQ64, tmdO64, tmQ128, tmKV64, tmdO128, stream);
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
