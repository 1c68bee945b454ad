// GPU-resident Hungarian matcher (see matcher.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sam3b {

struct MatcherArgs {
  int B = 0, Q = 0, Tmax = 0, repeats = 1;
  const float* logits = nullptr;        // [B][Q]      pred_logits.squeeze(-1)
  const float* pred_boxes = nullptr;    // [B][Q][4]   cxcywh
  const float* tgt_boxes = nullptr;     // [B][Tmax][4] cxcywh, padded
  const int32_t* num_boxes = nullptr;   // [B]
  const uint8_t* out_valid = nullptr;   // [B][Q] or null
  const uint8_t* tgt_valid = nullptr;   // [B][Tmax] or null
  float w_class = 1.f, w_bbox = 1.f, w_giou = 1.f;
  int focal = 0, stable = 0;
  float alpha = 0.25f, gamma = 2.f;
};

// cost [B][Q][Tmax] fp32 (written);  query_of_col [B][max(1,Tmax*repeats)], col_of_query [B][Q]: -1 = unmatched
int matcher_run(const MatcherArgs& a, float* cost, int32_t* query_of_col, int32_t* col_of_query, cudaStream_t s);

}  // namespace sam3b
