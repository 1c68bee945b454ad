// ViT trunk engine: owns the schedule of kernels for one forward / backward of the SAM3 image
// encoder trunk with LoRA adapters (reference: sam3/model/vitdet.py ViT.forward :813-859,
// Block.forward :597-613, Attention.forward :466-515; adapters lora_layers.py:49-55,87-91).
// Host-side C++ only; all device memory is provided by the caller (PyTorch owns it).
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "elementwise.cuh"

namespace sam3b {

enum LoraTarget : int { LT_Q = 1, LT_K = 2, LT_V = 4, LT_O = 8, LT_FC1 = 16, LT_FC2 = 32 };

struct VitConfig {
  int img_size = 1008, patch_size = 14, in_chans = 3, embed_dim = 1024, depth = 32, num_heads = 16;
  int mlp_hidden = 4736, window_size = 24;
  std::vector<int> global_blocks{7, 15, 23, 31};
  int pos_side = 24;  // pretrain grid side: pos_embed is [1][1 + side*side][D]
  float ln_eps = 1e-5f, rope_theta = 10000.f;
  int lora_rank = 16;
  float lora_scaling = 2.f;
  int lora_targets = LT_Q | LT_K | LT_V | LT_O | LT_FC1 | LT_FC2;
  int dtype = 0;  // tensor-core operand format: 0 fp16, 1 bf16
  int max_batch = 8;
};

struct LoraEntry {  // one adapter's place in the flat LoRA parameter / gradient buffers
  int block, target, in, out, rank;
  int64_t a_off, b_off;  // element offsets: A [in][rank], B [rank][out]
};

class VitEngine {
 public:
  explicit VitEngine(const VitConfig& cfg);
  const VitConfig& config() const { return cfg_; }
  int tokens_per_image() const { return G_ * G_; }
  int64_t weight_bytes() const { return weight_bytes_; }
  int64_t workspace_bytes(int batch, bool training);
  int64_t lora_numel() const { return lora_numel_; }
  const std::vector<LoraEntry>& lora_entries() const { return entries_; }
  static int num_base_tensors(int depth) { return 4 + 12 * depth; }

  // Binds caller-owned device memory. `batch`/`training` select the workspace layout (see workspace_bytes).
  int bind(void* weight_buf, int64_t weight_bytes, void* work_buf, int64_t work_bytes, int batch, bool training);
  // tensors (fp32, device), reference state-dict order:
  //   patch_embed.proj.weight, pos_embed, ln_pre.weight, ln_pre.bias, then per block:
  //   norm1.weight, norm1.bias, attn.qkv.weight, attn.qkv.bias, attn.proj.weight, attn.proj.bias,
  //   norm2.weight, norm2.bias, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias
  int load_base(const float* const* tensors, int n, cudaStream_t s);
  int forward(const float* img, int batch, const float* lora_flat, float* out_nchw, bool save, cudaStream_t s);
  int backward(const float* gout_nchw, float* lora_grad_flat, cudaStream_t s);
  // The same backward cut into block ranges [blk_hi .. blk_lo] (descending, consecutive, starting at depth - 1): each call
  // leaves the weight gradients of its blocks final in their slice of lora_grad_flat (lora_grad_range), so the caller can
  // all-reduce that slice while the next range runs.
  int backward_segment(const float* gout_nchw, float* lora_grad_flat, int blk_hi, int blk_lo, cudaStream_t s);
  void lora_grad_range(int blk_hi, int blk_lo, int64_t* lo, int64_t* hi) const;
  // Stochastic depth (DropPath, vitdet.py:610-611): device array [depth][2][batch] of per-image branch scales
  // (0 or 1/keep; [i][0] attention branch, [i][1] MLP branch) used by the next forward AND its backward.
  // nullptr disables it (eval mode / drop_path 0).
  void set_drop_path(const float* scales) { drop_scales_ = scales; }
  // Adapter dropout (lora_layers.py:43,54): probability and seed for the next forward and its backward; p = 0 disables.
  // seed_dev (optional): device word ADDED to `seed` inside the kernels, so a captured CUDA graph draws a new mask per replay
  // when the caller rewrites that word between replays.
  int set_lora_dropout(float p, uint32_t seed, const uint32_t* seed_dev = nullptr);

 private:
  struct Site {
    int in = 0, out = 0, n_ad = 0, R = 0, Rlive = 0;   // R: K-extension pitch (64-padded); Rlive: columns the K loop must cover
    int off[3] = {0, 0, 0}, len[3] = {0, 0, 0};
    int entry[3] = {-1, -1, -1};
    uint16_t *w_ext = nullptr, *wt_ext = nullptr, *down_T = nullptr, *up_pack = nullptr;
    int64_t ldw = 0, ldwt = 0;
    const float* bias = nullptr;
    float *dA_pack = nullptr, *dB_pack = nullptr;   // split-K weight-gradient accumulators of this site (training layouts)
  };
  struct BlockW {
    const float *g1, *b1, *g2, *b2;
    Site qkv, proj, fc1, fc2;
    bool global = false;
  };
  struct BlockAct {
    float *x_mid, *lse2, *mean1, *rstd1, *mean2, *rstd2;
    uint16_t *xn1, *qkv, *O, *xn2, *h, *g;
  };
  struct Bump;  // size-or-assign bump allocator (engine.cpp)
  void layout_weights(Bump& b);
  void layout_work(Bump& b, int batch, bool training);
  int pack_site(const Site& st, const float* lora_flat, cudaStream_t s) const;
  LoraSite make_site(const Site& st, const float* flat) const;
  int site_down(const Site& st, uint16_t* act, int64_t ld, int M, int block, int site, cudaStream_t s) const;
  int site_wgrad(const Site& st, const uint16_t* x_act, int64_t ldx, const uint16_t* dy_act, int64_t lddy, int M,
                 float* grad_flat, int block, int site, cudaStream_t s);

  VitConfig cfg_;
  int G_ = 0, T_ = 0, D_ = 0, H_ = 0, Dm_ = 0, Kpe_ = 0, Kpe_pad_ = 0;
  int64_t weight_bytes_ = 0, lora_numel_ = 0;
  std::vector<LoraEntry> entries_;
  std::vector<BlockW> blocks_;
  // device tables / packed weights
  float *rope_win_ = nullptr, *rope_glob_ = nullptr, *pos_tab_ = nullptr;
  float *ln_pre_g_ = nullptr, *ln_pre_b_ = nullptr;
  uint16_t* wpe_ = nullptr;
  float* small_ = nullptr;  // LN params and biases
  uint8_t* weight_base_ = nullptr;
  // workspace
  uint8_t* work_base_ = nullptr;
  int64_t work_bytes_ = 0;
  int bound_batch_ = 0;
  bool bound_training_ = false;
  uint16_t* patches_ = nullptr;
  std::vector<float*> x_;  // residual stream, x_[i] = input of block i, x_[depth] = output
  std::vector<BlockAct> acts_;
  float *dxa_ = nullptr, *dxb_ = nullptr, *delta_ = nullptr;
  // all sites' split-K accumulators are one region (zeroed by ONE memset per backward); descriptors of every adapted
  // Linear live in device memory so that packing / unpacking is one launch each per step (elementwise.cuh)
  float* wgrad_pack_ = nullptr;
  int64_t wgrad_pack_bytes_ = 0;
  LoraSiteDesc* site_desc_dev_ = nullptr;
  std::vector<LoraSiteDesc> site_desc_host_;
  int site_max_work_ = 0;
  std::vector<int> site_first_;    // first descriptor of each block (+ end sentinel)
  float *bwd_dx_ = nullptr, *bwd_dx_alt_ = nullptr;   // residual-gradient ping-pong state between backward segments
  int bwd_next_ = -1;              // next block a backward segment must start at
  bool site_desc_dirty_ = true;
  void build_site_descs();
  float* gscale_ = nullptr;   // device [s, 1/s, scratch]: power-of-two scale of the incoming gradient (grad_scale, conv.cuh)
  uint16_t *dx16_ = nullptr, *dh16_ = nullptr, *dxn16_ = nullptr, *dO16_ = nullptr, *dqkv16_ = nullptr;
  int Rmax_ = 0;
  float drop_p_ = 0.f, fwd_drop_p_ = 0.f;
  uint32_t drop_seed_ = 0, fwd_drop_seed_ = 0;
  const uint32_t *drop_seed_dev_ = nullptr, *fwd_drop_seed_dev_ = nullptr;
  uint16_t* xd16_ = nullptr;  // dropout(x) scratch [M][max in]
  const float* drop_scales_ = nullptr;
  const float* fwd_drop_scales_ = nullptr;  // what the saved forward used
  int last_batch_ = 0;
  bool last_saved_ = false;
  bool base_loaded_ = false;
};

}  // namespace sam3b
