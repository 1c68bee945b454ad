// FlashAttention-style forward for the SAM3 ViT (head_dim 64) on tcgen05 + TMA.
//
// Reference semantics: sam3/model/vitdet.py:502  x = F.scaled_dot_product_attention(q, k, v)
// (no mask, no dropout, scale = head_dim^-0.5) applied per window (576 tokens) or per image
// (5184 tokens).  RoPE (vitdet.py:485) has already been applied to q,k by the qkv GEMM epilogue;
// window_partition/unpartition (vitdet.py:93-139) do not exist here because the token stream
// is kept in window-major order, so a window is a contiguous run of L rows ("segment").
//
// Layout: qkv [tokens][ld] 16-bit with q at column h*64, k at D + h*64, v at 2D + h*64.
// One CTA = one (128-query tile, head, segment), 192 threads, FOUR CTAs co-resident per SM (48 KB smem, 128 TMEM
// columns and <= 85 registers each).  The exp2 of the softmax is the floor of this kernel (MUFU: 512 clk per
// 128x64 block per SM, tools/micro/mufu_rate.cu); one CTA alone keeps the MUFU pipe ~30 % busy because its
// MMA -> softmax -> MMA chain is serial, so the design goal is simply "as many independent chains per SM as TMEM allows":
//   warps 0-3 : softmax + output (thread t owns query row t == TMEM lane t)
//   warp  4   : TMA producer (Q once, then K_j / V_j into 2-stage rings)
//   warp  5   : MMA issuer + TMEM allocator; issue order S_0, PV_0, S_1, PV_1, ...
// Per 64-key block j:  S_j = Q K_j^T (SS MMA, fp32, TMEM columns [0,64)) -> two passes over the 64 scores, 32 at a time
// from TMEM (row max; then p = exp2(s*c - m*c), row sum) -> P_j (16-bit) written back over S_j's first 32 columns with
// tcgen05.st -> O += P_j V_j with P_j as the TMEM A operand (no shared-memory round trip, no proxy fence) and V_j as
// MN-major B.  The running max only moves when
// the block max exceeds it by more than 2^8 in the exp2 domain ("lazy rescale"): then the softmax
// warps rescale the O accumulator in TMEM (tcgen05.ld / st) before publishing P_j.  Output: O
// 16-bit, LSE in log2 units.
// History (profiles/r01_attn_bwd_timeline.md): v1 kept Q, P in shared memory (smem-bandwidth bound, 2 CTAs/SM, 542 TF/s
// on the global blocks); v2 moved Q and P to TMEM (646 TF/s); v3 (this) trades the S double buffer for a third and fourth CTA (window 0.183 ms = 534 TF/s, global 1.25 ms = 706 TF/s).
#include "attn.cuh"

#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace sam3b {

#ifdef SAM3B_TRACE
// Debug timeline (tools/attn_trace.py fwd): one chosen CTA stamps clock64() at its pipeline events.
__device__ unsigned long long g_attn_trace_fwd[4096];
#define TRACE(slot) do { if (blockIdx.x == 200 && blockIdx.y == 3 && (slot) < 4096) g_attn_trace_fwd[(slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;  // queries per CTA
constexpr int BKV = 64;  // keys per block: divides 576 and 5184
constexpr int Q_BYTES = BQ * HD * 2;   // 16 KB
constexpr int KV_BYTES = BKV * HD * 2; // 8 KB
#ifndef SAM3B_FWD_CTAS
#define SAM3B_FWD_CTAS 4               // co-resident CTAs per SM (TMEM: 128 columns each); measured 3 -> 4: window 0.211 -> 0.183 ms, global 1.43 -> 1.25 ms
#endif
constexpr int NS = SAM3B_FWD_CTAS >= 4 ? 2 : 3;   // K / V ring depth
// N CTAs/SM: N x (dynamic + 1 KB reserved) must fit the SM's 228 KB: no alignment slack, the dynamic
// window is declared 1024-byte aligned and checked at run time.
constexpr int FWD_SMEM = Q_BYTES + 2 * NS * KV_BYTES + 256 /*barriers*/;
constexpr int TCOLS = 128;  // S / P: [0,64)   O: [64,128)
constexpr float RESCALE_LOG2 = 8.f;  // p <= 2^8 between rescales: safe in fp16/bf16 and fp32 sums

struct FwdParams {
  int Lq, Lk;     // rows per segment (queries / keys)
  int q_tiles;    // ceil(Lq / 128)
  int q_col0, k_col0, v_col0, o_col0;
  const void* q; int64_t ldq;
  void* O; int64_t ldo;
  float* lse2;    // [H][nseg * Lq_stat]
  int Lq_stat;    // Lq rounded up to 64
  int64_t stat_stride;  // nseg * Lq_stat
  int H;
  float scale_log2;  // scale * log2(e)
  // GEN only
  const float* bias;     // [nseg*H][Lq][Lk] or null
  const uint8_t* kpm;    // [nseg][Lk] or null
  float drop_inv_keep; uint32_t drop_thr, drop_seed; const uint32_t* drop_bits; int bias_vec4; int bits_pitch_k;
};

// dropout mask on the attention probabilities: stateless hash of (seed, segment*H + head, query, key)
__device__ __forceinline__ uint32_t attn_drop_base(uint32_t seed, uint32_t bh) { return lowbias32(seed ^ (bh * 0x9E3779B1u + 0x85EBCA6Bu)); }
__device__ __forceinline__ bool attn_drop_keep(uint32_t base, uint32_t q, uint32_t k, uint32_t Lk, uint32_t thr) {
  return lowbias32(base ^ (q * Lk + k)) >= thr;
}

// GEN: additive mask / key padding / dropout (hash or keep-bits) / Lq != Lk.  DROPB (with GEN = false): the ViT instantiation
// plus dropout through precomputed keep-bits and nothing else - the encoder's 5184 x 5184 self-attention in training mode -
// so that dropout alone does not drag in the general path (log2-domain rescaling, per-half predicates, spills).
template <int DT, bool GEN, bool DROPB = false>
__global__ void __launch_bounds__(192, (GEN || DROPB) ? 3 : SAM3B_FWD_CTAS)   // the masked / dropout instantiations need ~95 registers
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // swizzle-128B operands need 1024-byte alignment
  uint8_t* sQ = smem_raw;
  uint8_t* sK = sQ + Q_BYTES;          // NS stages
  uint8_t* sV = sK + NS * KV_BYTES;    // NS stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * KV_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;            // [NS]
  uint64_t* k_free = k_full + NS;         // [NS]
  uint64_t* v_full = k_free + NS;         // [NS]
  uint64_t* v_free = v_full + NS;         // [NS]
  uint64_t* s_full = v_free + NS;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  static_assert((1 + 4 * NS + 3) * 8 + 4 <= 256, "barrier block overflows its 256 bytes");

  const int warp = threadIdx.x >> 5;
  const int q_tile = blockIdx.x % p.q_tiles;
  const int seg = blockIdx.x / p.q_tiles;
  const int head = blockIdx.y;
  const int kv_row0_seg = seg * p.Lk;
  const int q_row0 = seg * p.Lq + q_tile * BQ;
  const int n_blocks = (p.Lk + BKV - 1) / BKV;

  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_free[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_free[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();                             // after the TMEM allocation (see ptx.cuh)
  pdl_wait();                                // q / k / v come from the previous kernel
  const uint32_t tmem_S = *tmem_slot;        // S_j (fp32, 64 columns); P_j (16-bit) is written back over columns [0,32)
  const uint32_t tmem_O = tmem_S + BKV;

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, p.q_col0 + head * HD, q_row0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % NS;
        const int kv_row0 = kv_row0_seg + j * BKV;
        if (j >= NS) mbar_wait(&k_free[st], (j / NS - 1) & 1, 10);
        mbar_arrive_expect_tx(&k_full[st], KV_BYTES);
        tma_load_2d(sK + st * KV_BYTES, &tmKV, &k_full[st], p.k_col0 + head * HD, kv_row0);
        if (j >= NS) mbar_wait(&v_free[st], (j / NS - 1) & 1, 11);
        mbar_arrive_expect_tx(&v_full[st], KV_BYTES);
        tma_load_2d(sV + st * KV_BYTES, &tmKV, &v_full[st], p.v_col0 + head * HD, kv_row0);
      }
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(BQ, BKV, DT, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(BQ, HD, DT, 0, 1);
      constexpr uint32_t ST16 = KV_BYTES >> 4;
      const uint64_t dQ = make_desc_kmajor(smem_u32(sQ)), dK = make_desc_kmajor(smem_u32(sK));
      const uint64_t dV = make_desc_mnmajor(smem_u32(sV), 8192);
      // S_j = Q K_j^T, both operands in shared memory (48 clk per instruction instead of 32: irrelevant next to the
      // 512 clk of exp2 per block)
      auto issue_s = [&](int j) {
        const int st = j % NS;
        mbar_wait(&k_full[st], (j / NS) & 1, 21);
        tc_fence_after();
        const uint64_t kd = dK + (uint64_t)(st * ST16);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem_S, dQ + k * 2, kd + k * 2, idesc_s, k > 0);
        umma_commit(s_full);
        umma_commit(&k_free[st]);
        if (j < 96) TRACE(1024 + j * 4 + 0);
      };
      mbar_wait(q_full, 0, 20);
      issue_s(0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % NS;
        mbar_wait(p_full, j & 1, 22);
        mbar_wait(&v_full[st], (j / NS) & 1, 23);
        tc_fence_after();
        const uint64_t vd = dV + (uint64_t)(st * ST16);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)   // A = P_j in TMEM (16-bit, packed over the first 32 columns of S)
          umma_f16_ts(tmem_O, tmem_S + k * 8, vd + k * 128, idesc_pv, (j > 0 || k > 0));
        umma_commit(pv_done);
        umma_commit(&v_free[st]);
        if (j < 96) TRACE(1024 + j * 4 + 1);
        // S_{j+1} overwrites S_j / P_j: issued after P_j.V_j, and the tensor pipe executes in issue order
        if (j + 1 < n_blocks) issue_s(j + 1);
      }
    }
  } else {
    // ------------------------------ softmax / output ------------------------------
    const int r = threadIdx.x;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    const float tau = GEN ? RESCALE_LOG2 : RESCALE_LOG2 / c;  // in the units of the running max
    float m_used = -INFINITY, l_run = 0.f;
    // GEN: scores are first moved to the log2 domain (t = s*c + bias*log2e, -inf where the key is padded) and the
    // running max lives in that domain (c_eff = 1); the ViT instantiation keeps raw scores and folds c into the exp.
    const int q_in_seg = min(q_tile * BQ + r, p.Lq - 1);
    const float* bias_row = nullptr;
    uint32_t drop_base = 0;
    const uint32_t* bits_row = nullptr;      // precomputed keep-bits of this query row (one word per 32 keys)
    if constexpr (GEN) {
      if (p.bias != nullptr) bias_row = p.bias + ((int64_t)(seg * p.H + head) * p.Lq + q_in_seg) * p.Lk;
      drop_base = attn_drop_base(p.drop_seed, (uint32_t)(seg * p.H + head));
    }
    if constexpr (GEN || DROPB) {
      if (p.drop_bits != nullptr) bits_row = p.drop_bits + ((int64_t)(seg * p.H + head) * p.Lq + q_in_seg) * p.bits_pitch_k;
    }
    const float c_eff = GEN ? 1.f : c;

    for (int j = 0; j < n_blocks; ++j) {
      const int kv_valid = min(BKV, p.Lk - j * BKV);
      const bool full = kv_valid == BKV;   // predicate-free path (every block when Lk % 64 == 0)
      const bool tr = threadIdx.x == 0 && j < 96;
      if (tr) TRACE(64 + j * 8 + 0);
      mbar_wait(s_full, j & 1, 30);
      tc_fence_after();
      if (tr) TRACE(64 + j * 8 + 1);
      // 32 scores of this row (half h of the block) in the units the running max uses.  The block is read twice from
      // TMEM (max pass, exp pass) instead of being held in 64 registers: four CTAs per SM leave 85 registers per thread.
      auto load_half = [&](int h, float (&t)[32]) {
        uint32_t a[32];
        tmem_ld_x32(tmem_S + lane_off + h * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = __uint_as_float(a[i]);
        if constexpr (GEN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] *= c;
          if (bias_row != nullptr) {
            const float* bp = bias_row + j * BKV + h * 32;
            if (full && p.bias_vec4) {
              // every lane reads its OWN row (a gather of 32 lines per instruction): 16-byte loads need a quarter of the
              // LSU wavefronts of scalar ones (row a7, 401 x 5184 decoder cross-attention: the bias is 532 MB per call)
              const float4* b4 = reinterpret_cast<const float4*>(bp);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 b = __ldg(b4 + q4);
                t[4 * q4] = fmaf(b.x, 1.4426950408889634f, t[4 * q4]);
                t[4 * q4 + 1] = fmaf(b.y, 1.4426950408889634f, t[4 * q4 + 1]);
                t[4 * q4 + 2] = fmaf(b.z, 1.4426950408889634f, t[4 * q4 + 2]);
                t[4 * q4 + 3] = fmaf(b.w, 1.4426950408889634f, t[4 * q4 + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (h * 32 + i < kv_valid) t[i] = fmaf(__ldg(bp + i), 1.4426950408889634f, t[i]);
            }
          }
          if (p.kpm != nullptr) {
            const uint8_t* kp = p.kpm + (int64_t)seg * p.Lk + j * BKV + h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (h * 32 + i < kv_valid && __ldg(kp + i) != 0) t[i] = -INFINITY;
          }
        }
      };
      // ---- pass 1: block max ----
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float t[32];
        load_half(h, t);
        if (full) {
          float mx[4] = {t[0], t[1], t[2], t[3]};   // independent chains
#pragma unroll
          for (int i = 4; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], t[i]);
          m_blk = fmaxf(m_blk, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (h * 32 + i < kv_valid) m_blk = fmaxf(m_blk, t[i]);
        }
      }
      if (tr) TRACE(64 + j * 8 + 2);
      // always true on the first block (m_used = -inf) unless every key so far is masked (m_blk = -inf)
      const bool need = m_blk > m_used + tau;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? m_blk : m_used;
        const float alpha = ex2_approx((m_used - m_new) * c_eff);  // 0 on the first block, 1 for rows that keep their max
        if (j > 0) {
          // s_full(j) was signalled by a commit issued after P.V of block j-1: every earlier P.V has landed in O
#pragma unroll
          for (int cc = 0; cc < HD; cc += 32) {
            uint32_t t[32];
            tmem_ld_x32(tmem_O + lane_off + cc, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st_x32(tmem_O + lane_off + cc, t);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        m_used = m_new;
      }
      if (tr) TRACE(64 + j * 8 + 3);
      // ---- pass 2: p = exp2(t - m), row sum, P_j (16-bit) over the S_j columns ----
      const float mc = (m_used == -INFINITY) ? 0.f : m_used * c_eff;  // all keys masked so far: exp2(-inf - 0) = 0
      float l_part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float t[32];
        load_half(h, t);
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float e[8];
          if (full) {
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(t[q * 8 + i], c_eff, -mc));
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float v = ex2_approx(fmaf(t[q * 8 + i], c_eff, -mc));
              e[i] = (h * 32 + q * 8 + i < kv_valid) ? v : 0.f;
            }
          }
          // fp32 row sum of the unrounded, un-dropped probabilities
          l_part[q] += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          if constexpr (DROPB && !GEN) {     // keep-bits only (the launcher guarantees bits_row != nullptr)
            const uint32_t mw = __ldg(bits_row + (j * (BKV / 32) + h));
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = ((mw >> (q * 8 + i)) & 1u) ? e[i] * p.drop_inv_keep : 0.f;
          }
          if constexpr (GEN) {
            if (p.drop_thr != 0) {   // dropout on the probabilities fed to P.V
              if (bits_row != nullptr) {
                const int wi_ = j * (BKV / 32) + h;
                const uint32_t mw = wi_ < p.bits_pitch_k ? __ldg(bits_row + wi_) : 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = ((mw >> (q * 8 + i)) & 1u) ? e[i] * p.drop_inv_keep : 0.f;
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  e[i] = attn_drop_keep(drop_base, (uint32_t)q_in_seg, (uint32_t)(j * BKV + h * 32 + q * 8 + i), (uint32_t)p.Lk, p.drop_thr)
                             ? e[i] * p.drop_inv_keep : 0.f;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[q * 4 + i] = pack2<DT>(e[2 * i], e[2 * i + 1]);
        }
        // half 0 lands on S columns [0,16), half 1 on [16,32): both already consumed by this thread
        tmem_st_x16(tmem_S + lane_off + h * 16, pk);
      }
      l_run += (l_part[0] + l_part[1]) + (l_part[2] + l_part[3]);
      if (tr) TRACE(64 + j * 8 + 4);
      tmem_st_wait();
      tc_fence_before();         // our tcgen05.ld/st are ordered before the MMA warp's next tcgen05 ops
      mbar_arrive(p_full);
      if (tr) TRACE(64 + j * 8 + 5);
    }
    mbar_wait(pv_done, (n_blocks - 1) & 1, 33);
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;  // a fully masked row yields zeros (torch would give NaN)
    const int row = q_row0 + r;
    const bool valid = (q_tile * BQ + r) < p.Lq;
#pragma unroll
    for (int cc = 0; cc < HD; cc += 32) {
      uint32_t t[32];
      tmem_ld_x32(tmem_O + lane_off + cc, t);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.O) + (int64_t)row * p.ldo + p.o_col0 + head * HD + cc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack2<DT>(__uint_as_float(t[q * 8 + 0]) * inv_l, __uint_as_float(t[q * 8 + 1]) * inv_l);
          u.y = pack2<DT>(__uint_as_float(t[q * 8 + 2]) * inv_l, __uint_as_float(t[q * 8 + 3]) * inv_l);
          u.z = pack2<DT>(__uint_as_float(t[q * 8 + 4]) * inv_l, __uint_as_float(t[q * 8 + 5]) * inv_l);
          u.w = pack2<DT>(__uint_as_float(t[q * 8 + 6]) * inv_l, __uint_as_float(t[q * 8 + 7]) * inv_l);
          dst[q] = u;
        }
      }
    }
    if (valid)   // [heads][nseg * Lq_stat]
      p.lse2[(int64_t)head * p.stat_stride + (int64_t)seg * p.Lq_stat + q_tile * BQ + r] = m_used * c_eff + log2f(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_S, TCOLS);
}

// Keep-bits of the attention-dropout mask, both orientations, from the SAME hash the kernels evaluate inline.  One CTA per
// 256 x 256 (query, key) tile: warp w owns query rows 32w .. 32w+31 (lane = row) and walks the eight 32-key groups, so a lane
// ends up with 8 row words = 32 contiguous bytes of `bits`; a ballot per key gives the transposed word, collected in shared
// memory so that thread t then writes the 8 words = 32 bytes of key t's row of `bitsT`.  Row pitches are padded to 8 words
// (attn_bits_pitch) so every store is an aligned, full 32-byte sector.  Hashing every score once here instead of once in each
// of the three kernels is what makes dropout on a 5184 x 5184 attention affordable (bench.py, row a7).
__global__ void __launch_bounds__(256) attn_dropout_bits_kernel(int n_bh, int Lq, int Lk, int pitch_k, int pitch_q, uint32_t seed,
                                                                uint32_t thr, uint32_t* __restrict__ bits, uint32_t* __restrict__ bitsT) {
  __shared__ uint32_t colw[256][9];      // [key in tile][query group], padded against bank conflicts
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qt_n = (Lq + 255) / 256, kt_n = (Lk + 255) / 256;
  const int64_t n_tiles = (int64_t)n_bh * qt_n * kt_n;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int kt = (int)(t % kt_n), qt = (int)((t / kt_n) % qt_n), bh = (int)(t / ((int64_t)kt_n * qt_n));
    const uint32_t base = attn_drop_base(seed, (uint32_t)bh);
    const int q = qt * 256 + warp * 32 + lane;
    uint32_t row[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint32_t r = 0, col = 0;
      const int k0 = kt * 256 + g * 32;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const bool keep = attn_drop_keep(base, (uint32_t)q, (uint32_t)(k0 + k), (uint32_t)Lk, thr);
        r |= (keep ? 1u : 0u) << k;
        const uint32_t b = __ballot_sync(0xffffffffu, keep);      // bit l = keep(query row of lane l, key k0 + k)
        if (lane == k) col = b;
      }
      row[g] = r;
      colw[g * 32 + lane][warp] = col;
    }
    if (q < Lq) {
      uint4* dst = reinterpret_cast<uint4*>(bits + ((int64_t)bh * Lq + q) * pitch_k + kt * 8);
      dst[0] = make_uint4(row[0], row[1], row[2], row[3]);
      dst[1] = make_uint4(row[4], row[5], row[6], row[7]);
    }
    __syncthreads();
    const int key = kt * 256 + threadIdx.x;
    if (key < Lk) {
      const uint32_t* c = colw[threadIdx.x];
      uint4* dst = reinterpret_cast<uint4*>(bitsT + ((int64_t)bh * Lk + key) * pitch_q + qt * 8);
      dst[0] = make_uint4(c[0], c[1], c[2], c[3]);
      dst[1] = make_uint4(c[4], c[5], c[6], c[7]);
    }
    __syncthreads();
  }
}

// one instantiation (and one cached smem attribute) per (operand format, feature set)
template <int DT, bool GEN, bool DROPB = false>
static int launch_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const FwdParams& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DT, GEN, DROPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr_set = true;
  }
  SAM3B_CHECK_CUDA(launch_pdl(attn_fwd_kernel<DT, GEN, DROPB>, grid, dim3(192), FWD_SMEM, stream, tmQ, tmKV, p));
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace

int attn_fwd_launch(const AttnArgs& a, cudaStream_t stream) {
  SAM3B_REQUIRE(a.q && a.kv && a.O && a.lse2, "attention: null tensor");
  SAM3B_REQUIRE(a.nseg > 0 && a.Lq > 0 && a.Lk > 0 && a.heads > 0, "attention: empty problem");
  SAM3B_REQUIRE(a.ldo % 8 == 0 && a.ldq % 8 == 0 && a.ldkv % 8 == 0, "attention: leading dimensions must be multiples of 8");
  SAM3B_REQUIRE(a.q_col0 % 8 == 0 && a.k_col0 % 8 == 0 && a.v_col0 % 8 == 0 && a.o_col0 % 8 == 0, "attention: column offsets must be multiples of 8");
  SAM3B_REQUIRE(a.dtype == 0 || a.dtype == 1, "attention: dtype");
  SAM3B_REQUIRE(a.drop_p >= 0.f && a.drop_p < 1.f, "attention: dropout p=%f outside [0,1)", a.drop_p);
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_2d(&tmQ, a.q, (uint64_t)a.nseg * a.Lq, a.q_cols, a.ldq, BQ, HD);
  if (rc) return rc;
  rc = make_tmap_2d(&tmKV, a.kv, (uint64_t)a.nseg * a.Lk, a.kv_cols, a.ldkv, BKV, HD);
  if (rc) return rc;
  FwdParams p{};
  p.Lq = a.Lq; p.Lk = a.Lk; p.q_tiles = (a.Lq + BQ - 1) / BQ;
  p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0; p.o_col0 = a.o_col0;
  p.q = a.q; p.ldq = a.ldq;
  p.O = a.O; p.ldo = a.ldo; p.lse2 = a.lse2; p.H = a.heads;
  p.Lq_stat = attn_lq_stat(a.Lq); p.stat_stride = (int64_t)a.nseg * p.Lq_stat;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.bias = a.bias; p.kpm = a.kpm;
  p.drop_inv_keep = 1.f / (1.f - a.drop_p); p.drop_thr = a.drop_p > 0.f ? dropout_threshold(a.drop_p) : 0u; p.drop_seed = a.drop_seed; p.drop_bits = (a.drop_p > 0.f && a.Lq % 32 == 0 && a.Lk % 32 == 0) ? a.drop_bits : nullptr;
  p.bits_pitch_k = attn_bits_pitch(a.Lk);
  p.bias_vec4 = (a.bias != nullptr && a.Lk % 4 == 0 && (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) ? 1 : 0;
  dim3 grid(p.q_tiles * a.nseg, a.heads);
  const bool gen = a.bias != nullptr || a.kpm != nullptr || a.drop_p > 0.f;
  // dropout through keep-bits and nothing else, full blocks only: the ViT instantiation + the bit mask
  if (a.bias == nullptr && a.kpm == nullptr && a.drop_p > 0.f && p.drop_bits != nullptr && a.Lq == a.Lk && a.Lk % BKV == 0)
    return a.dtype == 0 ? launch_fwd<0, false, true>(tmQ, tmKV, p, grid, stream) : launch_fwd<1, false, true>(tmQ, tmKV, p, grid, stream);
  if (gen) return a.dtype == 0 ? launch_fwd<0, true>(tmQ, tmKV, p, grid, stream) : launch_fwd<1, true>(tmQ, tmKV, p, grid, stream);
  return a.dtype == 0 ? launch_fwd<0, false>(tmQ, tmKV, p, grid, stream) : launch_fwd<1, false>(tmQ, tmKV, p, grid, stream);
}

#ifdef SAM3B_TRACE
int attn_trace_read_fwd(unsigned long long* host, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_attn_trace_fwd, sizeof(unsigned long long) * (size_t)n);
}
#endif

int attn_bits_pitch(int L) { return ((L + 31) / 32 + 7) / 8 * 8; }

int attn_dropout_bits(int n_bh, int Lq, int Lk, float p, uint32_t seed, uint32_t* bits, uint32_t* bitsT, cudaStream_t stream) {
  SAM3B_REQUIRE(n_bh > 0 && Lq > 0 && Lk > 0 && Lq % 32 == 0 && Lk % 32 == 0, "attn_dropout_bits: Lq and Lk must be multiples of 32 (got %d, %d)", Lq, Lk);
  SAM3B_REQUIRE(p > 0.f && p < 1.f && bits && bitsT, "attn_dropout_bits: bad arguments");
  SAM3B_REQUIRE((reinterpret_cast<uintptr_t>(bits) & 31) == 0 && (reinterpret_cast<uintptr_t>(bitsT) & 31) == 0, "attn_dropout_bits: buffers must be 32-byte aligned");
  const int64_t tiles = (int64_t)n_bh * ((Lq + 255) / 256) * ((Lk + 255) / 256);
  const int blocks = (int)std::min<int64_t>(tiles, (int64_t)num_sms() * 8);
  attn_dropout_bits_kernel<<<blocks, 256, 0, stream>>>(n_bh, Lq, Lk, attn_bits_pitch(Lk), attn_bits_pitch(Lq), seed, dropout_threshold(p), bits, bitsT);
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace sam3b
