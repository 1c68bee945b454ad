// ViT trunk engine (see engine.h).  Everything here is host-side scheduling: it decides which
// kernel runs on which buffer; all arithmetic lives in gemm.cu / attn_*.cu / elementwise.cu.
//
// Data layout in HBM (per rank, batch B, M = B * T tokens, T = (img/patch)^2):
//   * token order is window-major: token = ((b*nwin + window)*ws + i)*ws + j, so a window is a
//     contiguous run of ws*ws rows and an image a run of T rows.  window_partition/unpartition
//     (vitdet.py:93-139) disappear; only the RoPE / position tables are stored in that order.
//   * residual stream x_[i] : fp32 [M][D] per block boundary (kept for the LayerNorm backward).
//   * GEMM operands are 16-bit (fp16 or bf16) with a K-extension: an adapted Linear reads
//     [act | s*(act.A)] (width in + R) against [W | B^T] (width in + R), R = 64-padded total rank,
//     so the LoRA up-projection is accumulated by the same tcgen05 K loop as the frozen weight.
//     The backward mirrors this with [dy | s*(dy.B^T)] against [W^T | A].
//   * frozen weights are stored twice (W and W^T, both K-major) so every GEMM on the hot path
//     uses the same K-major/K-major kernel; only the LoRA weight-gradients use MN-major operands.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "attn.cuh"
#include "common.h"
#include "conv.cuh"
#include "gemm.cuh"
#include "rng.cuh"

namespace sam3b {

struct VitEngine::Bump {
  uint8_t* base;
  int64_t off = 0;
  explicit Bump(uint8_t* b) : base(b) {}
  template <typename T>
  T* take(int64_t count) {
    off = (off + 1023) & ~int64_t(1023);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * (int64_t)sizeof(T);
    return p;
  }
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

VitEngine::VitEngine(const VitConfig& cfg) : cfg_(cfg) {
  G_ = cfg.img_size / cfg.patch_size;
  T_ = G_ * G_;
  D_ = cfg.embed_dim;
  H_ = cfg.num_heads;
  Dm_ = cfg.mlp_hidden;
  Kpe_ = cfg.in_chans * cfg.patch_size * cfg.patch_size;
  Kpe_pad_ = round_up(Kpe_, 64);
  blocks_.resize(cfg.depth);
  const int r = cfg.lora_rank;
  int64_t off = 0;
  auto add_entry = [&](int blk, int target, int in, int out) {
    LoraEntry e{blk, target, in, out, r, off, off + (int64_t)in * r};
    off += (int64_t)in * r + (int64_t)r * out;
    entries_.push_back(e);
    return (int)entries_.size() - 1;
  };
  for (int i = 0; i < cfg.depth; ++i) {
    BlockW& b = blocks_[i];
    b.global = std::find(cfg.global_blocks.begin(), cfg.global_blocks.end(), i) != cfg.global_blocks.end();
    b.qkv.in = D_; b.qkv.out = 3 * D_;
    b.proj.in = D_; b.proj.out = D_;
    b.fc1.in = D_; b.fc1.out = Dm_;
    b.fc2.in = Dm_; b.fc2.out = D_;
    const int qkv_bits[3] = {LT_Q, LT_K, LT_V};
    for (int t = 0; t < 3; ++t)
      if (r > 0 && (cfg.lora_targets & qkv_bits[t])) {
        Site& s = b.qkv;
        s.off[s.n_ad] = t * D_; s.len[s.n_ad] = D_;
        s.entry[s.n_ad] = add_entry(i, qkv_bits[t], D_, D_);
        ++s.n_ad;
      }
    auto single = [&](Site& s, int bit) {
      if (r > 0 && (cfg.lora_targets & bit)) {
        s.off[0] = 0; s.len[0] = s.out; s.entry[0] = add_entry(i, bit, s.in, s.out); s.n_ad = 1;
      }
    };
    single(b.proj, LT_O);
    single(b.fc1, LT_FC1);
    single(b.fc2, LT_FC2);
    for (Site* s : {&b.qkv, &b.proj, &b.fc1, &b.fc2}) {
      // K-extension width: total adapter rank rounded up to a whole 64-wide k-block.  Rounding to the MMA K step (16) would
      // skip 3 of the last k-block's 4 MMAs, but makes the operand row pitch (in + R) * 2 B a non-multiple of 128 B: every TMA
      // box row then straddles two lines; measured 0.5 % SLOWER per step (round-1 A/B), so 64 stays.
      s->R = s->n_ad > 0 ? round_up(s->n_ad * r, 64) : 0;
      // ... but the K loop only has to cover the live columns: the operand pitch keeps the 64-wide pad (alignment), K stops at
      // the last 16-wide MMA step that holds adapter columns (the rest of the pad is zero in both operands)
      s->Rlive = s->n_ad > 0 ? round_up(s->n_ad * r, 16) : 0;
      s->ldw = s->in + s->R;
      s->ldwt = s->out + s->R;
      Rmax_ = std::max(Rmax_, s->R);
    }
  }
  lora_numel_ = off;
  Bump b(nullptr);
  layout_weights(b);
  weight_bytes_ = b.off + 1024;
}

void VitEngine::layout_weights(Bump& b) {
  const int ws2 = cfg_.window_size * cfg_.window_size;
  rope_win_ = b.take<float>((int64_t)ws2 * 64);
  rope_glob_ = b.take<float>((int64_t)T_ * 64);
  pos_tab_ = b.take<float>((int64_t)T_ * D_);
  wpe_ = b.take<uint16_t>((int64_t)D_ * Kpe_pad_);
  ln_pre_g_ = b.take<float>(D_);
  ln_pre_b_ = b.take<float>(D_);
  for (BlockW& w : blocks_) {
    float* g1 = b.take<float>(D_); float* b1 = b.take<float>(D_);
    float* g2 = b.take<float>(D_); float* b2 = b.take<float>(D_);
    w.g1 = g1; w.b1 = b1; w.g2 = g2; w.b2 = b2;
    for (Site* s : {&w.qkv, &w.proj, &w.fc1, &w.fc2}) {
      s->bias = b.take<float>(s->out);
      s->w_ext = b.take<uint16_t>((int64_t)s->out * s->ldw);
      s->wt_ext = b.take<uint16_t>((int64_t)s->in * s->ldwt);
      if (s->R > 0) {
        s->down_T = b.take<uint16_t>((int64_t)s->R * s->in);
        s->up_pack = b.take<uint16_t>((int64_t)s->R * s->out);
      }
    }
  }
}

// rows of the head-major attention statistics (lse2, delta): every segment of L queries is padded to a multiple of 64
static int64_t g_stat_rows(int64_t M, int L) { return (M / L) * (int64_t)((L + 63) / 64 * 64); }

void VitEngine::layout_work(Bump& b, int batch, bool training) {
  const int64_t M = (int64_t)batch * T_;
  const int64_t stat_rows = std::max(g_stat_rows(M, cfg_.window_size * cfg_.window_size), g_stat_rows(M, T_));
  const int depth = cfg_.depth;
  const int nsave = training ? depth : 1;
  patches_ = b.take<uint16_t>(M * Kpe_pad_);
  x_.assign(depth + 1, nullptr);
  if (training) {
    for (int i = 0; i <= depth; ++i) x_[i] = b.take<float>(M * D_);
  } else {
    float* p0 = b.take<float>(M * D_);
    float* p1 = b.take<float>(M * D_);
    for (int i = 0; i <= depth; ++i) x_[i] = (i & 1) ? p1 : p0;
  }
  acts_.assign(depth, BlockAct{});
  std::vector<BlockAct> saved(nsave);
  for (int i = 0; i < nsave; ++i) {
    const BlockW& w = blocks_[std::min(i, depth - 1)];
    BlockAct& a = saved[i];
    a.x_mid = b.take<float>(M * D_);
    a.lse2 = b.take<float>(stat_rows * H_);
    a.mean1 = b.take<float>(M); a.rstd1 = b.take<float>(M);
    a.mean2 = b.take<float>(M); a.rstd2 = b.take<float>(M);
    // widths use Rmax_ so that a single saved slot (inference) fits every block
    (void)w;
    a.xn1 = b.take<uint16_t>(M * (D_ + Rmax_));
    a.qkv = b.take<uint16_t>(M * (3 * D_));
    a.O = b.take<uint16_t>(M * (D_ + Rmax_));
    a.xn2 = b.take<uint16_t>(M * (D_ + Rmax_));
    a.h = b.take<uint16_t>(M * Dm_);
    a.g = b.take<uint16_t>(M * (Dm_ + Rmax_));
  }
  for (int i = 0; i < depth; ++i) acts_[i] = saved[training ? i : 0];
  if (training) {
    dxa_ = b.take<float>(M * D_);
    dxb_ = b.take<float>(M * D_);
    delta_ = b.take<float>(stat_rows * H_);
    dx16_ = b.take<uint16_t>(M * (D_ + Rmax_));
    dh16_ = b.take<uint16_t>(M * (Dm_ + Rmax_));
    dxn16_ = b.take<uint16_t>(M * D_);
    dO16_ = b.take<uint16_t>(M * D_);
    dqkv16_ = b.take<uint16_t>(M * (3 * D_ + Rmax_));
    xd16_ = b.take<uint16_t>(M * std::max(Dm_, D_));
    gscale_ = b.take<float>(4);
    // per-site split-K accumulators, contiguous
    int64_t total = 0;
    for (auto& w : blocks_)
      for (Site* st : {&w.qkv, &w.proj, &w.fc1, &w.fc2})
        if (st->R > 0) total += (int64_t)st->in * st->R + (int64_t)st->R * st->out;
    wgrad_pack_ = b.take<float>(std::max<int64_t>(total, 1));
    wgrad_pack_bytes_ = total * 4;
    float* cur = wgrad_pack_;
    for (auto& w : blocks_)
      for (Site* st : {&w.qkv, &w.proj, &w.fc1, &w.fc2}) {
        st->dA_pack = st->dB_pack = nullptr;
        if (st->R > 0) {
          st->dA_pack = cur; cur += (int64_t)st->in * st->R;
          st->dB_pack = cur; cur += (int64_t)st->R * st->out;
        }
      }
  } else {
    wgrad_pack_ = nullptr; wgrad_pack_bytes_ = 0;
    for (auto& w : blocks_)
      for (Site* st : {&w.qkv, &w.proj, &w.fc1, &w.fc2}) st->dA_pack = st->dB_pack = nullptr;
  }
  site_desc_dev_ = b.take<LoraSiteDesc>(4 * (int64_t)blocks_.size());
}

int64_t VitEngine::workspace_bytes(int batch, bool training) {
  // sizing pass on a scratch copy of the pointer members (restored by the next bind())
  Bump b(nullptr);
  layout_work(b, batch, training);
  if (work_base_ != nullptr) {  // re-establish the bound layout
    Bump r(work_base_);
    layout_work(r, bound_batch_, bound_training_);
  }
  return b.off + 1024;
}

int VitEngine::bind(void* weight_buf, int64_t weight_bytes, void* work_buf, int64_t work_bytes, int batch, bool training) {
  SAM3B_REQUIRE(weight_buf && work_buf, "vit bind: null buffer");
  SAM3B_REQUIRE((reinterpret_cast<uintptr_t>(weight_buf) & 1023) == 0 && (reinterpret_cast<uintptr_t>(work_buf) & 1023) == 0,
                "vit bind: buffers must be 1024-byte aligned");
  SAM3B_REQUIRE(weight_bytes >= weight_bytes_, "vit bind: weight buffer %lld < %lld bytes", (long long)weight_bytes, (long long)weight_bytes_);
  SAM3B_REQUIRE(batch >= 1 && batch <= cfg_.max_batch, "vit bind: batch %d outside [1, %d]", batch, cfg_.max_batch);
  if (weight_base_ != weight_buf) base_loaded_ = false;
  weight_base_ = static_cast<uint8_t*>(weight_buf);
  Bump bw(weight_base_);
  layout_weights(bw);
  Bump sz(nullptr);
  layout_work(sz, batch, training);
  SAM3B_REQUIRE(work_bytes >= sz.off, "vit bind: workspace %lld < %lld bytes", (long long)work_bytes, (long long)sz.off);
  work_base_ = static_cast<uint8_t*>(work_buf);
  work_bytes_ = work_bytes;
  bound_batch_ = batch;
  bound_training_ = training;
  Bump b(work_base_);
  layout_work(b, batch, training);
  build_site_descs();
  last_saved_ = false;
  return 0;
}

// ---------------------------------------------------------------------------------------------
int VitEngine::load_base(const float* const* t, int n, cudaStream_t s) {
  SAM3B_REQUIRE(weight_base_ != nullptr, "vit load_base: bind() first");
  SAM3B_REQUIRE(n == num_base_tensors(cfg_.depth), "vit load_base: expected %d tensors, got %d", num_base_tensors(cfg_.depth), n);
  for (int i = 0; i < n; ++i) SAM3B_REQUIRE(t[i] != nullptr, "vit load_base: tensor %d is null", i);
  const int dt = cfg_.dtype;
  // rope tables (float64 on the host; compute_axial_cis, vitdet.py:41-57), stored in window-major order
  {
    const int hd = D_ / H_, nf = hd / 4, ws = cfg_.window_size, nwx = G_ / ws;
    SAM3B_REQUIRE(hd == 64, "vit: head_dim %d unsupported (64 only)", hd);
    std::vector<double> freqs(nf);
    for (int m = 0; m < nf; ++m) freqs[m] = 1.0 / std::pow((double)cfg_.rope_theta, (4.0 * m) / hd);
    std::vector<float> win((size_t)ws * ws * 64), glob((size_t)T_ * 64);
    for (int tkn = 0; tkn < ws * ws; ++tkn) {
      const double tx = tkn % ws, ty = tkn / ws;
      for (int m = 0; m < nf; ++m) {
        win[(size_t)tkn * 64 + 2 * m] = (float)std::cos(tx * freqs[m]);
        win[(size_t)tkn * 64 + 2 * m + 1] = (float)std::sin(tx * freqs[m]);
        win[(size_t)tkn * 64 + 2 * (nf + m)] = (float)std::cos(ty * freqs[m]);
        win[(size_t)tkn * 64 + 2 * (nf + m) + 1] = (float)std::sin(ty * freqs[m]);
      }
    }
    const double sc = (double)ws / G_;  // rope_interp: scale_pos = rope_pt_size / input_size (vitdet.py:438-441)
    for (int tkn = 0; tkn < T_; ++tkn) {
      const int w = tkn / (ws * ws), in = tkn % (ws * ws);
      const int pi = (w / nwx) * ws + in / ws, pj = (w % nwx) * ws + in % ws;
      const double tx = pj * sc, ty = pi * sc;
      for (int m = 0; m < nf; ++m) {
        glob[(size_t)tkn * 64 + 2 * m] = (float)std::cos(tx * freqs[m]);
        glob[(size_t)tkn * 64 + 2 * m + 1] = (float)std::sin(tx * freqs[m]);
        glob[(size_t)tkn * 64 + 2 * (nf + m)] = (float)std::cos(ty * freqs[m]);
        glob[(size_t)tkn * 64 + 2 * (nf + m) + 1] = (float)std::sin(ty * freqs[m]);
      }
    }
    SAM3B_CHECK_CUDA(cudaStreamSynchronize(s));
    SAM3B_CHECK_CUDA(cudaMemcpy(rope_win_, win.data(), win.size() * 4, cudaMemcpyHostToDevice));
    SAM3B_CHECK_CUDA(cudaMemcpy(rope_glob_, glob.data(), glob.size() * 4, cudaMemcpyHostToDevice));
  }
  int rc;
  // patch embed weight [D][C*P*P] -> 16-bit [D][Kpe_pad] (zero pad)
  SAM3B_CHECK_CUDA(cudaMemsetAsync(wpe_, 0, (size_t)D_ * Kpe_pad_ * 2, s));
  if ((rc = pack_weight(t[0], D_, Kpe_, wpe_, Kpe_pad_, 0, dt, s))) return rc;
  if ((rc = build_pos_table(t[1], cfg_.pos_side, G_, cfg_.window_size, D_, pos_tab_, s))) return rc;
  auto copyf = [&](const float* dst, const float* src, int64_t n_) -> int {
    SAM3B_CHECK_CUDA(cudaMemcpyAsync(const_cast<float*>(dst), src, n_ * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
  };
  if ((rc = copyf(ln_pre_g_, t[2], D_))) return rc;
  if ((rc = copyf(ln_pre_b_, t[3], D_))) return rc;
  for (int i = 0; i < cfg_.depth; ++i) {
    const float* const* p = t + 4 + 12 * i;
    BlockW& w = blocks_[i];
    if ((rc = copyf(w.g1, p[0], D_)) || (rc = copyf(w.b1, p[1], D_)) || (rc = copyf(w.g2, p[6], D_)) ||
        (rc = copyf(w.b2, p[7], D_)))
      return rc;
    struct { Site* s; const float* W; const float* b; } sites[4] = {
        {&w.qkv, p[2], p[3]}, {&w.proj, p[4], p[5]}, {&w.fc1, p[8], p[9]}, {&w.fc2, p[10], p[11]}};
    for (auto& e : sites) {
      Site& st = *e.s;
      if ((rc = copyf(st.bias, e.b, st.out))) return rc;
      // zero the K-extensions, then the frozen parts
      SAM3B_CHECK_CUDA(cudaMemsetAsync(st.w_ext, 0, (size_t)st.out * st.ldw * 2, s));
      SAM3B_CHECK_CUDA(cudaMemsetAsync(st.wt_ext, 0, (size_t)st.in * st.ldwt * 2, s));
      if ((rc = pack_weight(e.W, st.out, st.in, st.w_ext, st.ldw, 0, dt, s))) return rc;
      if ((rc = pack_weight(e.W, st.out, st.in, st.wt_ext, st.ldwt, 1, dt, s))) return rc;
    }
  }
  base_loaded_ = true;
  return 0;
}

LoraSite VitEngine::make_site(const Site& st, const float* flat) const {
  LoraSite ls;
  ls.in = st.in; ls.out_total = st.out; ls.n = st.n_ad; ls.r = cfg_.lora_rank; ls.rpad = st.R;
  for (int a = 0; a < st.n_ad; ++a) {
    const LoraEntry& e = entries_[st.entry[a]];
    ls.out_off[a] = st.off[a]; ls.out_len[a] = st.len[a];
    ls.A[a] = flat + e.a_off; ls.B[a] = flat + e.b_off;
  }
  return ls;
}

int VitEngine::pack_site(const Site& st, const float* lora_flat, cudaStream_t s) const {
  if (st.R == 0) return 0;
  return lora_pack(make_site(st, lora_flat), st.down_T, st.w_ext, st.ldw, st.up_pack, st.wt_ext, st.ldwt, cfg_.dtype, s);
}

// Host copy of the device-resident descriptors of every adapted Linear (uploaded by the next forward).
void VitEngine::build_site_descs() {
  site_desc_host_.clear();
  site_max_work_ = 0;
  site_first_.assign(blocks_.size() + 1, 0);
  int blk_i = 0;
  for (auto& w : blocks_) {
    site_first_[blk_i++] = (int)site_desc_host_.size();
    for (const Site* st : {&w.qkv, &w.proj, &w.fc1, &w.fc2}) {
      if (st->R == 0) continue;
      LoraSiteDesc d{};
      d.in = st->in; d.out_total = st->out; d.n = st->n_ad; d.r = cfg_.lora_rank; d.rpad = st->R;
      for (int a = 0; a < st->n_ad; ++a) {
        const LoraEntry& e = entries_[st->entry[a]];
        d.out_off[a] = st->off[a]; d.out_len[a] = st->len[a];
        d.a_off[a] = e.a_off; d.b_off[a] = e.b_off;
      }
      d.down_T = st->down_T; d.w_ext = st->w_ext; d.up_pack = st->up_pack; d.wt_ext = st->wt_ext;
      d.ldw = st->ldw; d.ldwt = st->ldwt;
      d.dA_pack = st->dA_pack; d.dB_pack = st->dB_pack;
      site_desc_host_.push_back(d);
      site_max_work_ = std::max(site_max_work_, st->R * std::max(st->in, st->out));
    }
  }
  site_first_[blocks_.size()] = (int)site_desc_host_.size();
  site_desc_dirty_ = true;
}

int VitEngine::set_lora_dropout(float p, uint32_t seed, const uint32_t* seed_dev) {
  SAM3B_REQUIRE(p >= 0.f && p < 1.f, "vit set_lora_dropout: p=%f outside [0,1)", p);
  drop_p_ = p;
  drop_seed_ = seed;
  drop_seed_dev_ = seed_dev;
  return 0;
}

// act[:, in .. in+R) = s * drop(act[:, :in]) . A   (skinny GEMM; B operand = down_T [R][in]).  With adapter
// dropout the A operand is the masked copy xd16_ (the base branch keeps reading the unmasked act).
int VitEngine::site_down(const Site& st, uint16_t* act, int64_t ld, int M, int block, int site, cudaStream_t s) const {
  if (st.R == 0) return 0;
  const uint16_t* src = act;
  int64_t lds = ld;
  if (drop_p_ > 0.f) {
    SAM3B_REQUIRE(xd16_ != nullptr, "vit: adapter dropout needs a training workspace");
    int rc = dropout_rows16(act, ld, M, st.in, xd16_, st.in, drop_p_, site_seed(drop_seed_, block, site), cfg_.dtype, s, drop_seed_dev_);
    if (rc) return rc;
    src = xd16_;
    lds = st.in;
  }
  GemmArgs a;
  a.M = M; a.N = st.R; a.K = st.in;
  a.A = src; a.lda = lds; a.B = st.down_T; a.ldb = st.in;
  a.dtype = cfg_.dtype; a.epilogue = EPI_STORE16; a.alpha = cfg_.lora_scaling;
  a.C = act + st.in; a.ldc = ld; a.bn = 64;
  return gemm_launch(a, s);
}

// LoRA weight gradients of one site.  x_act: forward input [M][in | T'] (T' = s*x.A in the extension),
// dy_act: [M][out | dT''] (dT'' = s*dy.B^T in the extension).
//   dB_a = T'_a^T . dy[:, slice_a]      dA_a = x^T . dT''_a
int VitEngine::site_wgrad(const Site& st, const uint16_t* x_act, int64_t ldx, const uint16_t* dy_act, int64_t lddy, int M,
                          float* grad_flat, int block, int site, cudaStream_t s) {
  if (st.R == 0) return 0;
  const uint16_t* xa = x_act;   // operand of dA = drop(x)^T . dT''
  int64_t ldxa = ldx;
  if (fwd_drop_p_ > 0.f) {
    int rcd = dropout_rows16(x_act, ldx, M, st.in, xd16_, st.in, fwd_drop_p_, site_seed(fwd_drop_seed_, block, site), cfg_.dtype, s, fwd_drop_seed_dev_);
    if (rcd) return rcd;
    xa = xd16_;
    ldxa = st.in;
  }
  SAM3B_REQUIRE(st.dA_pack && st.dB_pack, "vit backward: no weight-gradient workspace (bind with training = 1)");
  const int kb_total = (M + 63) / 64;
  auto splitk_for = [&](int feat) {
    const int m_tiles = (feat + 127) / 128;
    int sk = std::max(1, num_sms() / m_tiles);
    return std::max(1, std::min(sk, kb_total / 4));
  };
  int rc;
  {  // dB_pack [R][out] (transposed store of C[out][nr]) = dy^T . T'
    GemmArgs a;
    a.M = st.out; a.N = st.R; a.K = M;
    a.A = dy_act; a.lda = lddy; a.a_mn = 1;
    a.B = x_act + st.in; a.ldb = ldx; a.b_mn = 1;
    a.dtype = cfg_.dtype; a.epilogue = EPI_ATOMIC_F32;
    a.C = st.dB_pack; a.ldc = st.out; a.c_trans = 1; a.splitk = splitk_for(st.out);
    if ((rc = gemm_launch(a, s))) return rc;
  }
  {  // dA_pack [in][R] = x^T . dT''
    GemmArgs a;
    a.M = st.in; a.N = st.R; a.K = M;
    a.A = xa; a.lda = ldxa; a.a_mn = 1;
    a.B = dy_act + st.out; a.ldb = lddy; a.b_mn = 1;
    a.dtype = cfg_.dtype; a.epilogue = EPI_ATOMIC_F32;
    a.C = st.dA_pack; a.ldc = st.R; a.splitk = splitk_for(st.in);
    if ((rc = gemm_launch(a, s))) return rc;
  }
  (void)grad_flat;   // un-packed for all sites at once at the end of backward() (lora_unpack_all)
  return 0;
}

// ---------------------------------------------------------------------------------------------
int VitEngine::forward(const float* img, int batch, const float* lora_flat, float* out_nchw, bool save, cudaStream_t s) {
  SAM3B_REQUIRE(base_loaded_, "vit forward: load_base() first");
  SAM3B_REQUIRE(batch >= 1 && batch <= bound_batch_, "vit forward: batch %d exceeds bound batch %d", batch, bound_batch_);
  SAM3B_REQUIRE(!save || bound_training_, "vit forward: save_for_backward needs a training workspace");
  SAM3B_REQUIRE(lora_numel_ == 0 || lora_flat != nullptr, "vit forward: lora parameters missing");
  SAM3B_REQUIRE(drop_p_ == 0.f || bound_training_, "vit forward: adapter dropout needs a training workspace");
  const int M = batch * T_;
  const int dt = cfg_.dtype;
  const int ws2 = cfg_.window_size * cfg_.window_size;
  int rc;
  // patch embed: gather -> GEMM (+ tiled abs pos) -> ln_pre (vitdet.py:814-833)
  if ((rc = patch_gather(img, batch, cfg_.in_chans, cfg_.img_size, cfg_.img_size, cfg_.patch_size, cfg_.window_size,
                         patches_, Kpe_pad_, Kpe_pad_, dt, s)))
    return rc;
  {
    GemmArgs a;
    a.M = M; a.N = D_; a.K = Kpe_pad_;
    a.A = patches_; a.lda = Kpe_pad_; a.B = wpe_; a.ldb = Kpe_pad_;
    a.dtype = dt; a.epilogue = EPI_RESIDUAL_F32;
    a.residual = pos_tab_; a.ldres = D_; a.res_row_mod = T_;
    a.C = x_[0]; a.ldc = D_;
    if ((rc = gemm_launch(a, s))) return rc;
  }
  if ((rc = layernorm_fwd_f32(x_[0], ln_pre_g_, ln_pre_b_, cfg_.ln_eps, M, D_, x_[0], s))) return rc;

  // adapter operands of ALL sites in one launch (the descriptors are uploaded once per bind, outside any graph capture)
  if (!site_desc_host_.empty()) {
    if (site_desc_dirty_) {
      SAM3B_CHECK_CUDA(cudaMemcpyAsync(site_desc_dev_, site_desc_host_.data(), site_desc_host_.size() * sizeof(LoraSiteDesc),
                                       cudaMemcpyHostToDevice, s));
      site_desc_dirty_ = false;
    }
    if ((rc = lora_pack_all(site_desc_dev_, (int)site_desc_host_.size(), site_max_work_, lora_flat, dt, s))) return rc;
  }

  for (int i = 0; i < cfg_.depth; ++i) {
    BlockW& w = blocks_[i];
    BlockAct& a = acts_[i];
    const int64_t ld_xn1 = D_ + w.qkv.R, ld_O = D_ + w.proj.R, ld_xn2 = D_ + w.fc1.R, ld_g = Dm_ + w.fc2.R;
    // ---- attention half: x_mid = x + proj(attn(rope(qkv(LN1(x)))))
    if ((rc = layernorm_fwd(x_[i], w.g1, w.b1, cfg_.ln_eps, M, D_, a.xn1, ld_xn1, dt, a.mean1, a.rstd1, s))) return rc;
    if ((rc = site_down(w.qkv, a.xn1, ld_xn1, M, i, 0, s))) return rc;
    {
      GemmArgs g;
      g.M = M; g.N = 3 * D_; g.K = D_ + w.qkv.Rlive;
      g.A = a.xn1; g.lda = ld_xn1; g.B = w.qkv.w_ext; g.ldb = w.qkv.ldw;
      g.dtype = dt; g.epilogue = EPI_QKV_ROPE; g.bias = w.qkv.bias;
      g.rope = w.global ? rope_glob_ : rope_win_; g.rope_period = w.global ? T_ : ws2; g.rope_cols = 2 * D_;
      g.C = a.qkv; g.ldc = 3 * D_;
      if ((rc = gemm_launch(g, s))) return rc;
    }
    {
      const int L = w.global ? T_ : ws2;
      AttnArgs f;
      f.q = a.qkv; f.ldq = 3 * D_; f.q_cols = 3 * D_; f.q_col0 = 0;
      f.kv = a.qkv; f.ldkv = 3 * D_; f.kv_cols = 3 * D_; f.k_col0 = D_; f.v_col0 = 2 * D_;
      f.nseg = M / L; f.Lq = L; f.Lk = L; f.heads = H_; f.scale = 0.125f; f.dtype = dt;
      f.O = a.O; f.ldo = ld_O; f.lse2 = a.lse2;
      if ((rc = attn_fwd_launch(f, s))) return rc;
    }
    if ((rc = site_down(w.proj, a.O, ld_O, M, i, 1, s))) return rc;
    {
      GemmArgs g;
      g.M = M; g.N = D_; g.K = D_ + w.proj.Rlive;
      g.A = a.O; g.lda = ld_O; g.B = w.proj.w_ext; g.ldb = w.proj.ldw;
      g.dtype = dt; g.epilogue = EPI_RESIDUAL_F32; g.bias = w.proj.bias;
      g.residual = x_[i]; g.ldres = D_;
      if (drop_scales_) { g.row_scale = drop_scales_ + (int64_t)(2 * i) * batch; g.rows_per_scale = T_; }
      g.C = a.x_mid; g.ldc = D_;
      if ((rc = gemm_launch(g, s))) return rc;
    }
    // ---- MLP half: x_out = x_mid + fc2(gelu(fc1(LN2(x_mid))))
    if ((rc = layernorm_fwd(a.x_mid, w.g2, w.b2, cfg_.ln_eps, M, D_, a.xn2, ld_xn2, dt, a.mean2, a.rstd2, s))) return rc;
    if ((rc = site_down(w.fc1, a.xn2, ld_xn2, M, i, 2, s))) return rc;
    {
      GemmArgs g;
      g.M = M; g.N = Dm_; g.K = D_ + w.fc1.Rlive;
      g.A = a.xn2; g.lda = ld_xn2; g.B = w.fc1.w_ext; g.ldb = w.fc1.ldw;
      g.dtype = dt; g.epilogue = EPI_GELU; g.bias = w.fc1.bias;
      g.C = a.h; g.ldc = Dm_; g.C2 = a.g; g.ldc2 = ld_g;
      if ((rc = gemm_launch(g, s))) return rc;
    }
    if ((rc = site_down(w.fc2, a.g, ld_g, M, i, 3, s))) return rc;
    {
      GemmArgs g;
      g.M = M; g.N = D_; g.K = Dm_ + w.fc2.Rlive;
      g.A = a.g; g.lda = ld_g; g.B = w.fc2.w_ext; g.ldb = w.fc2.ldw;
      g.dtype = dt; g.epilogue = EPI_RESIDUAL_F32; g.bias = w.fc2.bias;
      g.residual = a.x_mid; g.ldres = D_;
      if (drop_scales_) { g.row_scale = drop_scales_ + (int64_t)(2 * i + 1) * batch; g.rows_per_scale = T_; }
      g.C = x_[i + 1]; g.ldc = D_;
      if ((rc = gemm_launch(g, s))) return rc;
    }
  }
  if (out_nchw != nullptr)
    if ((rc = tokens_to_nchw(x_[cfg_.depth], batch, G_, cfg_.window_size, D_, out_nchw, s))) return rc;
  last_batch_ = batch;
  last_saved_ = save;
  fwd_drop_scales_ = drop_scales_;
  fwd_drop_p_ = drop_p_;
  fwd_drop_seed_ = drop_seed_;
  fwd_drop_seed_dev_ = drop_seed_dev_;
  return 0;
}

// ---------------------------------------------------------------------------------------------
int VitEngine::backward(const float* gout_nchw, float* grad_flat, cudaStream_t s) {
  return backward_segment(gout_nchw, grad_flat, cfg_.depth - 1, 0, s);
}

// Blocks [blk_hi .. blk_lo] of the backward (descending).  The first segment (blk_hi = depth - 1) consumes gout; the
// residual-gradient ping-pong buffers carry the state to the next segment; every segment ends by un-packing the weight
// gradients of ITS blocks into grad_flat (a contiguous slice, lora_grad_range), so a data-parallel caller can start the
// all-reduce of that slice while the next segment runs (reference DDP semantics: native_trainer.py:322-340).
int VitEngine::backward_segment(const float* gout_nchw, float* grad_flat, int blk_hi, int blk_lo, cudaStream_t s) {
  SAM3B_REQUIRE(last_saved_, "vit backward: needs a forward(save_for_backward=1) first");
  SAM3B_REQUIRE(lora_numel_ == 0 || grad_flat, "vit backward: null argument");
  SAM3B_REQUIRE(blk_hi < cfg_.depth && blk_lo >= 0 && blk_lo <= blk_hi, "vit backward: bad block range [%d, %d]", blk_hi, blk_lo);
  const bool first = blk_hi == cfg_.depth - 1;
  SAM3B_REQUIRE(first ? gout_nchw != nullptr : bwd_next_ == blk_hi,
                "vit backward: segment [%d, %d] out of order (next expected block %d)", blk_hi, blk_lo, bwd_next_);
  const int M = last_batch_ * T_;
  const int dt = cfg_.dtype;
  const int ws2 = cfg_.window_size * cfg_.window_size;
  int rc;
  int64_t ld_dx16 = D_ + Rmax_;
  const float* ds = fwd_drop_scales_;
  const int Bn = last_batch_;
  auto drop = [&](int blk, int branch) -> const float* { return ds ? ds + (int64_t)(2 * blk + branch) * Bn : nullptr; };
  if (first) {
    bwd_dx_ = dxa_;       // gradient w.r.t. the current block's output (fp32)
    bwd_dx_alt_ = dxb_;
    // The whole backward is linear in gout, so it runs on s * gout with s a power of two chosen on the device from
    // max|gout| (fp16 operands would otherwise flush a mean-reduced loss gradient, ~1e-9 per element, to zero); the
    // LoRA gradients are multiplied by 1/s when they are unpacked into grad_flat.
    if ((rc = grad_scale(gout_nchw, (int64_t)last_batch_ * D_ * T_, 256.f, gscale_, s))) return rc;
    if (wgrad_pack_bytes_ > 0) SAM3B_CHECK_CUDA(cudaMemsetAsync(wgrad_pack_, 0, (size_t)wgrad_pack_bytes_, s));   // all split-K accumulators at once
    if ((rc = nchw_to_tokens(gout_nchw, last_batch_, G_, cfg_.window_size, D_, bwd_dx_, dx16_, ld_dx16, dt, s, drop(cfg_.depth - 1, 1), gscale_))) return rc;
  }
  float*& dx = bwd_dx_;
  float*& dx_alt = bwd_dx_alt_;

  // dst16[:, out .. out+R) = s * dy[:, :out] . B^T      (skinny GEMM; B operand = up_pack [R][out])
  auto site_up_grad = [&](const Site& st, uint16_t* dy, int64_t ld) -> int {
    if (st.R == 0) return 0;
    GemmArgs a;
    a.M = M; a.N = st.R; a.K = st.out;
    a.A = dy; a.lda = ld; a.B = st.up_pack; a.ldb = st.out;
    a.dtype = dt; a.epilogue = EPI_STORE16; a.alpha = cfg_.lora_scaling;
    a.C = dy + st.out; a.ldc = ld; a.bn = 64;
    return gemm_launch(a, s);
  };
  // dst = [dy | dT''] . [W^T | A]^T  (dgrad through the frozen weight + the adapter in one K loop)
  auto site_dgrad = [&](const Site& st, const uint16_t* dy, int64_t ld, int epi, void* dst, int64_t lddst,
                        const void* aux, int64_t ldaux, int block, int site) -> int {
    const bool masked = fwd_drop_p_ > 0.f && st.R > 0;
    GemmArgs a;
    a.M = M; a.N = st.in; a.K = masked ? st.out : st.out + st.Rlive;
    a.A = dy; a.lda = ld; a.B = st.wt_ext; a.ldb = st.ldwt;
    a.dtype = dt; a.epilogue = epi; a.C = dst; a.ldc = lddst; a.aux = aux; a.ldaux = ldaux;
    int r2 = gemm_launch(a, s);
    if (r2 || !masked) return r2;
    // adapter part under dropout: dst += mask/(1-p) * (dT'' . A^T) [* gelu'(h)], K = R
    GemmArgs b;
    b.M = M; b.N = st.in; b.K = st.Rlive;
    b.A = dy + st.out; b.lda = ld; b.B = st.wt_ext + st.out; b.ldb = st.ldwt;
    b.dtype = dt; b.epilogue = EPI_ADDMASK16; b.C = dst; b.ldc = lddst;
    if (epi == EPI_DGELU) { b.aux = aux; b.ldaux = ldaux; }
    b.drop_p = fwd_drop_p_; b.drop_seed = site_seed(fwd_drop_seed_, block, site); b.drop_seed_dev = fwd_drop_seed_dev_;
    return gemm_launch(b, s);
  };

  for (int i = blk_hi; i >= blk_lo; --i) {
    BlockW& w = blocks_[i];
    BlockAct& a = acts_[i];
    const int64_t ld_xn1 = D_ + w.qkv.R, ld_O = D_ + w.proj.R, ld_xn2 = D_ + w.fc1.R, ld_g = Dm_ + w.fc2.R;
    const int64_t ld_dh = Dm_ + Rmax_, ld_dqkv = 3 * D_ + Rmax_;
    // ---- MLP half.  dx16_ holds dy for fc2 (width D [+R]).
    if ((rc = site_up_grad(w.fc2, dx16_, ld_dx16))) return rc;
    if ((rc = site_wgrad(w.fc2, a.g, ld_g, dx16_, ld_dx16, M, grad_flat, i, 3, s))) return rc;
    if ((rc = site_dgrad(w.fc2, dx16_, ld_dx16, EPI_DGELU, dh16_, ld_dh, a.h, Dm_, i, 3))) return rc;
    if ((rc = site_up_grad(w.fc1, dh16_, ld_dh))) return rc;
    if ((rc = site_wgrad(w.fc1, a.xn2, ld_xn2, dh16_, ld_dh, M, grad_flat, i, 2, s))) return rc;
    if ((rc = site_dgrad(w.fc1, dh16_, ld_dh, EPI_STORE16, dxn16_, D_, nullptr, 0, i, 2))) return rc;
    // dx_mid = dx + dLN2(dxn2)
    if ((rc = layernorm_bwd(dxn16_, D_, a.x_mid, a.mean2, a.rstd2, w.g2, dx, M, D_, dx_alt, dx16_, ld_dx16, dt, s, drop(i, 0), T_))) return rc;
    std::swap(dx, dx_alt);
    // ---- attention half.  dx16_ holds dy for proj.
    if ((rc = site_up_grad(w.proj, dx16_, ld_dx16))) return rc;
    if ((rc = site_wgrad(w.proj, a.O, ld_O, dx16_, ld_dx16, M, grad_flat, i, 1, s))) return rc;
    if (fwd_drop_p_ > 0.f && w.proj.R > 0) {
      // adapter dropout splits this dgrad into two GEMMs: delta needs the final dO, so it stays a separate pass
      if ((rc = site_dgrad(w.proj, dx16_, ld_dx16, EPI_STORE16, dO16_, D_, nullptr, 0, i, 1))) return rc;
      if ((rc = attn_delta(dO16_, D_, a.O, ld_O, M, H_, dt, delta_, s))) return rc;
    } else {
      // dO = dy . [W^T | A]^T and delta = rowsum(dO * O) per head in the same epilogue (O read once, no extra launch)
      const int L = w.global ? T_ : ws2;
      SAM3B_CHECK_CUDA(cudaMemsetAsync(delta_, 0, (size_t)g_stat_rows(M, L) * H_ * sizeof(float), s));
      GemmArgs g;
      g.M = M; g.N = w.proj.in; g.K = w.proj.out + w.proj.Rlive;
      g.A = dx16_; g.lda = ld_dx16; g.B = w.proj.wt_ext; g.ldb = w.proj.ldwt;
      g.dtype = dt; g.epilogue = EPI_STORE16_DELTA; g.C = dO16_; g.ldc = D_;
      g.aux = a.O; g.ldaux = ld_O;
      g.delta = delta_; g.delta_Lq = L; g.delta_Lq_stat = (L + 63) / 64 * 64; g.delta_stride = (int64_t)(M / L) * g.delta_Lq_stat;
      if ((rc = gemm_launch(g, s))) return rc;
    }
    {
      const int L = w.global ? T_ : ws2;
      AttnArgs b;
      b.q = a.qkv; b.ldq = 3 * D_; b.q_cols = 3 * D_; b.q_col0 = 0;
      b.kv = a.qkv; b.ldkv = 3 * D_; b.kv_cols = 3 * D_; b.k_col0 = D_; b.v_col0 = 2 * D_;
      b.nseg = M / L; b.Lq = L; b.Lk = L; b.heads = H_; b.scale = 0.125f; b.dtype = dt;
      b.O = a.O; b.ldo = ld_O; b.lse2 = a.lse2;
      b.dO = dO16_; b.lddo = D_; b.delta = delta_;
      b.dq = dqkv16_; b.lddq = ld_dqkv; b.dq_col0 = 0;
      b.dkv = dqkv16_; b.lddkv = ld_dqkv; b.dk_col0 = D_; b.dv_col0 = 2 * D_;
      b.rope = w.global ? rope_glob_ : rope_win_; b.rope_period = L;
      if ((rc = attn_bwd_launch(b, s))) return rc;
    }
    if ((rc = site_up_grad(w.qkv, dqkv16_, ld_dqkv))) return rc;
    if ((rc = site_wgrad(w.qkv, a.xn1, ld_xn1, dqkv16_, ld_dqkv, M, grad_flat, i, 0, s))) return rc;
    if (i == 0) break;  // nothing upstream of block 0 is trainable (patch embed / pos / ln_pre are frozen)
    if ((rc = site_dgrad(w.qkv, dqkv16_, ld_dqkv, EPI_STORE16, dxn16_, D_, nullptr, 0, i, 0))) return rc;
    if ((rc = layernorm_bwd(dxn16_, D_, x_[i], a.mean1, a.rstd1, w.g1, dx, M, D_, dx_alt, dx16_, ld_dx16, dt, s, drop(i - 1, 1), T_))) return rc;
    std::swap(dx, dx_alt);
  }
  // this segment's packed weight gradients -> their slice of the flat gradient buffer, times 1/s, in one launch
  if (!site_desc_host_.empty()) {
    const int d0 = site_first_[blk_lo], d1 = site_first_[blk_hi + 1];
    if (d1 > d0 && (rc = lora_unpack_all(site_desc_dev_ + d0, d1 - d0, site_max_work_, grad_flat, gscale_ + 1, s))) return rc;
  }
  bwd_next_ = blk_lo - 1;
  if (blk_lo == 0) last_saved_ = false;
  return 0;
}

// Element range [lo, hi) of the flat LoRA buffers that belongs to blocks [blk_lo .. blk_hi].
void VitEngine::lora_grad_range(int blk_hi, int blk_lo, int64_t* lo, int64_t* hi) const {
  int64_t a = lora_numel_, b = 0;
  for (const LoraEntry& e : entries_)
    if (e.block >= blk_lo && e.block <= blk_hi) {
      a = std::min(a, e.a_off);
      b = std::max(b, e.b_off + (int64_t)e.rank * e.out);
    }
  *lo = std::min(a, b);
  *hi = b;
}

}  // namespace sam3b
