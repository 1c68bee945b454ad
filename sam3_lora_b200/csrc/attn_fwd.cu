// FlashAttention-style forward for the SAM3 ViT (head_dim 64) on tcgen05 + TMA.
//
// Reference semantics: sam3/model/vitdet.py:502  x = F.scaled_dot_product_attention(q, k, v)
// (no mask, no dropout, scale = head_dim^-0.5) applied per window (576 tokens) or per image
// (5184 tokens).  RoPE (vitdet.py:485) has already been applied to q,k by the qkv GEMM epilogue;
// window_partition/unpartition (vitdet.py:93-139) do not exist here because the token stream
// is kept in window-major order, so a window is a contiguous run of L rows ("segment").
//
// Layout: qkv [tokens][ld] 16-bit with q at column h*64, k at D + h*64, v at 2D + h*64.
// One CTA = one (128-query tile, head, segment); 2 CTAs are co-resident per SM (<= 113 KB smem,
// 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//   warps 0-3 : softmax + output (thread t owns query row t == TMEM lane t)
//   warp  4   : TMA producer (Q once, then K_j / V_j, single-buffered)
//   warp  5   : MMA issuer + TMEM allocator
// Per 192-key block j:  S = Q K_j^T (tcgen05, fp32 in TMEM) -> online softmax in registers ->
// P (16-bit) to 128B-swizzled smem -> PV_j = P V_j (V as MN-major B operand) -> accumulated into
// the fp32 register copy of O with the running rescale.  Output: O 16-bit, LSE in log2 units.
#include "attn.cuh"

#include "common.h"
#include "ptx.cuh"

namespace sam3b {

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;   // queries per CTA
constexpr int BKV = 192;  // keys per block: divides 576 and 5184
constexpr int Q_BYTES = BQ * HD * 2;    // 16 KB
constexpr int K_BYTES = BKV * HD * 2;   // 24 KB
constexpr int V_BYTES = BKV * HD * 2;   // 24 KB
constexpr int P_BYTES = BQ * BKV * 2;   // 48 KB = 3 swizzle atoms of [128][64]
// 2 CTAs/SM: 2 x (dynamic + 1 KB reserved) must fit the SM's 228 KB, so no alignment slack: the
// dynamic window is declared 1024-byte aligned and checked at run time.
constexpr int FWD_SMEM = Q_BYTES + K_BYTES + V_BYTES + P_BYTES + 128 /*barriers*/;
constexpr int TCOLS = 256;  // S: [0,192)  PV: [192,256)

struct FwdParams {
  int L;          // tokens per segment
  int q_tiles;    // ceil(L / 128)
  int D;          // model width (q/k/v column blocks are D apart)
  void* O; int64_t ldo;
  float* lse2;    // [tokens][H]
  int H;
  float scale_log2;  // head_dim^-0.5 * log2(e)
  int total_rows;
};

template <int DT>
__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // swizzle-128B operands need 1024-byte alignment
  uint8_t* sQ = smem_raw;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + K_BYTES;
  uint8_t* sP = sV + V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_free = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_free = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int q_tile = blockIdx.x % p.q_tiles;
  const int seg = blockIdx.x / p.q_tiles;
  const int head = blockIdx.y;
  const int seg_row0 = seg * p.L;
  const int q_row0 = seg_row0 + q_tile * BQ;
  const int n_blocks = (p.L + BKV - 1) / BKV;

  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_free, 1); mbar_init(v_full, 1); mbar_init(v_free, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_PV = tmem_base + BKV;

  if (warp == 4) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, head * HD, q_row0);
      for (int j = 0; j < n_blocks; ++j) {
        const int kv_row0 = seg_row0 + j * BKV;
        if (j > 0) mbar_wait(k_free, (j - 1) & 1, 10);
        mbar_arrive_expect_tx(k_full, K_BYTES);
        tma_load_2d(sK, &tmKV, k_full, p.D + head * HD, kv_row0);
        if (j > 0) mbar_wait(v_free, (j - 1) & 1, 11);
        mbar_arrive_expect_tx(v_full, V_BYTES);
        tma_load_2d(sV, &tmKV, v_full, 2 * p.D + head * HD, kv_row0);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(BQ, BKV, DT, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(BQ, HD, DT, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      mbar_wait(q_full, 0, 20);
      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait(k_full, j & 1, 21);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_f16_ss(tmem_S, make_desc_kmajor(q_addr + k * 32), make_desc_kmajor(k_addr + k * 32), idesc_s, k > 0);
        umma_commit(s_full);
        umma_commit(k_free);
        mbar_wait(p_full, j & 1, 22);
        mbar_wait(v_full, j & 1, 23);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_f16_ss(tmem_PV, make_desc_kmajor(p_addr + (k >> 2) * (BQ * 128) + (k & 3) * 32),
                      make_desc_mnmajor(v_addr + k * 2048, 8192), idesc_pv, k > 0);
        umma_commit(o_full);
        umma_commit(v_free);
      }
    }
  } else {
    // ------------------------------ softmax / output ------------------------------
    const int r = threadIdx.x;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const float c = p.scale_log2;
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < n_blocks; ++j) {
      const int kv_valid = min(BKV, p.L - j * BKV);
      mbar_wait(s_full, j & 1, 30);
      tc_fence_after();
      // pass 1: row max over the valid keys of this block
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < BKV; cc += 32) {
        uint32_t s[32];
        tmem_ld_x32(tmem_S + lane_off + cc, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cc + i < kv_valid) m_blk = fmaxf(m_blk, __uint_as_float(s[i]));
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = exp2f((m_run - m_new) * c);  // exp2(-inf) = 0 on the first block
      // fold in the previous block's P.V, then rescale
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1, 31);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < HD; cc += 32) {
          uint32_t t[32];
          tmem_ld_x32(tmem_PV + lane_off + cc, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[cc + i] = (o_acc[cc + i] + __uint_as_float(t[i])) * alpha;
        }
      }
      l_run *= alpha;
      m_run = m_new;
      // pass 2: p = exp2(s*c - m*c), write 16-bit P into the swizzled A-operand tile
      const float mc = m_new * c;
      float l_blk = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < BKV; cc += 32) {
        uint32_t s[32];
        tmem_ld_x32(tmem_S + lane_off + cc, s);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = exp2f(__uint_as_float(s[i]) * c - mc);
          pv[i] = (cc + i < kv_valid) ? e : 0.f;
        }
        uint8_t* atom = p_row + (cc >> 6) * (BQ * 128);
        const int ch0 = (cc & 63) >> 3;  // first 16-byte chunk of this 32-column group inside the atom row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack2<DT>(pv[q * 8 + 0], pv[q * 8 + 1]);
          u.y = pack2<DT>(pv[q * 8 + 2], pv[q * 8 + 3]);
          u.z = pack2<DT>(pv[q * 8 + 4], pv[q * 8 + 5]);
          u.w = pack2<DT>(pv[q * 8 + 6], pv[q * 8 + 7]);
          // the row sum uses the rounded probabilities that the P.V MMA will actually see
          float2 f0 = unpack2<DT>(u.x), f1 = unpack2<DT>(u.y), f2 = unpack2<DT>(u.z), f3 = unpack2<DT>(u.w);
          l_blk += (f0.x + f0.y) + (f1.x + f1.y) + (f2.x + f2.y) + (f3.x + f3.y);
          *reinterpret_cast<uint4*>(atom + (((ch0 + q) ^ sw) << 4)) = u;
        }
      }
      l_run += l_blk;
      tc_fence_before();       // our tcgen05.ld of S complete before the MMA warp overwrites S
      fence_proxy_async_smem();  // st.shared of P visible to the tensor core (async proxy)
      mbar_arrive(p_full);
    }
    // last block's P.V
    mbar_wait(o_full, (n_blocks - 1) & 1, 32);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int row = q_row0 + r;
    const bool valid = (q_tile * BQ + r) < p.L && row < p.total_rows;
#pragma unroll
    for (int cc = 0; cc < HD; cc += 32) {
      uint32_t t[32];
      tmem_ld_x32(tmem_PV + lane_off + cc, t);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.O) + (int64_t)row * p.ldo + head * HD + cc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack2<DT>((o_acc[cc + q * 8 + 0] + __uint_as_float(t[q * 8 + 0])) * inv_l,
                          (o_acc[cc + q * 8 + 1] + __uint_as_float(t[q * 8 + 1])) * inv_l);
          u.y = pack2<DT>((o_acc[cc + q * 8 + 2] + __uint_as_float(t[q * 8 + 2])) * inv_l,
                          (o_acc[cc + q * 8 + 3] + __uint_as_float(t[q * 8 + 3])) * inv_l);
          u.z = pack2<DT>((o_acc[cc + q * 8 + 4] + __uint_as_float(t[q * 8 + 4])) * inv_l,
                          (o_acc[cc + q * 8 + 5] + __uint_as_float(t[q * 8 + 5])) * inv_l);
          u.w = pack2<DT>((o_acc[cc + q * 8 + 6] + __uint_as_float(t[q * 8 + 6])) * inv_l,
                          (o_acc[cc + q * 8 + 7] + __uint_as_float(t[q * 8 + 7])) * inv_l);
          dst[q] = u;
        }
      }
    }
    if (valid) p.lse2[(int64_t)row * p.H + head] = m_run * c + log2f(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, TCOLS);
}

// one instantiation (and one cached smem attribute) per operand format
template <int DT>
static int launch_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const FwdParams& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr_set = true;
  }
  attn_fwd_kernel<DT><<<grid, 192, FWD_SMEM, stream>>>(tmQ, tmKV, p);
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace

int attn_fwd_launch(const AttnFwdArgs& a, cudaStream_t stream) {
  SAM3B_REQUIRE(a.head_dim == 64, "attention: head_dim %d not supported (64 only)", a.head_dim);
  SAM3B_REQUIRE(a.tokens % a.seg_len == 0, "attention: tokens %d not a multiple of seg_len %d", a.tokens, a.seg_len);
  SAM3B_REQUIRE(a.heads * 64 == a.D, "attention: heads*64 != D");
  SAM3B_REQUIRE(a.ldo % 8 == 0 && a.ldqkv % 8 == 0, "attention: leading dimensions must be multiples of 8");
  SAM3B_REQUIRE(a.dtype == 0 || a.dtype == 1, "attention: dtype");
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_2d(&tmQ, a.qkv, a.tokens, 3 * a.D, a.ldqkv, BQ, HD);
  if (rc) return rc;
  rc = make_tmap_2d(&tmKV, a.qkv, a.tokens, 3 * a.D, a.ldqkv, BKV, HD);
  if (rc) return rc;
  FwdParams p{};
  p.L = a.seg_len; p.q_tiles = (a.seg_len + BQ - 1) / BQ; p.D = a.D;
  p.O = a.O; p.ldo = a.ldo; p.lse2 = a.lse2; p.H = a.heads;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.total_rows = a.tokens;
  const int nseg = a.tokens / a.seg_len;
  dim3 grid(p.q_tiles * nseg, a.heads);
  return a.dtype == 0 ? launch_fwd<0>(tmQ, tmKV, p, grid, stream) : launch_fwd<1>(tmQ, tmKV, p, grid, stream);
}

}  // namespace sam3b
