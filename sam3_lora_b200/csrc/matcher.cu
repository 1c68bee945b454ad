// GPU-resident Hungarian matcher (next-row f2): replaces `BinaryHungarianMatcherV2.forward` + `_do_matching`
// (sam3/train/matcher.py:15-29, 431-668), i.e. the cost matrix built by ~15 torch kernels, its .cpu().numpy() copy and one
// scipy.optimize.linear_sum_assignment call per image, by two kernels and no device->host transfer.
//
//   matcher_cost_kernel   C[b][q][t] = w_bbox * L1(cxcywh) + w_class * cost_class + w_giou * (-GIoU)   (fp32, the
//                         reference's arithmetic; 1e9 where the prediction or the target is flagged invalid)
//   matcher_lsa_kernel    one CTA per image: shortest-augmenting-path assignment with dual variables (Crouse 2016, the
//                         algorithm behind scipy's rectangular_lsap) in float64 like scipy.  The smaller side of the
//                         (queries x targets*repeats) problem is the row side, exactly as scipy transposes.  The inner
//                         scan over the free columns is spread over the CTA's threads and closed by a block-wide arg-min.
// Integer / index work: the result is bit-exact with scipy whenever the optimum is unique (always, for float costs
// without exact ties).  Exact ties (tiled `repeats` columns, 1e9 entries) are broken towards a new sink first, like scipy;
// the matched (query, target) SET is then still the reference's.
#include "matcher.cuh"

#include <algorithm>
#include <cfloat>

#include "common.h"

namespace sam3b {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// torch.nn.functional.logsigmoid: min(x, 0) - log1p(exp(-|x|))
__device__ __forceinline__ float logsigmoidf_(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(256) matcher_cost_kernel(const MatcherArgs a, float* __restrict__ C) {
  const int64_t total = (int64_t)a.B * a.Q * a.Tmax;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % a.Tmax), q = (int)((i / a.Tmax) % a.Q), b = (int)(i / ((int64_t)a.Tmax * a.Q));
    const float* pb = a.pred_boxes + ((int64_t)b * a.Q + q) * 4;
    const float* tb = a.tgt_boxes + ((int64_t)b * a.Tmax + t) * 4;
    const float pcx = pb[0], pcy = pb[1], pw = pb[2], ph = pb[3];
    const float tcx = tb[0], tcy = tb[1], tw = tb[2], th = tb[3];
    const float cost_bbox = ((fabsf(pcx - tcx) + fabsf(pcy - tcy)) + fabsf(pw - tw)) + fabsf(ph - th);   // torch.cdist(p=1)
    // box_cxcywh_to_xyxy + generalized_box_iou (sam3/model/box_ops.py:11-14, 91-142)
    const float px0 = pcx - 0.5f * pw, py0 = pcy - 0.5f * ph, px1 = pcx + 0.5f * pw, py1 = pcy + 0.5f * ph;
    const float tx0 = tcx - 0.5f * tw, ty0 = tcy - 0.5f * th, tx1 = tcx + 0.5f * tw, ty1 = tcy + 0.5f * th;
    const float area1 = (px1 - px0) * (py1 - py0), area2 = (tx1 - tx0) * (ty1 - ty0);
    const float iw = fmaxf(fminf(px1, tx1) - fmaxf(px0, tx0), 0.f), ih = fmaxf(fminf(py1, ty1) - fmaxf(py0, ty0), 0.f);
    const float inter = iw * ih;
    const float uni = area1 + area2 - inter;
    const float iou = inter / uni;
    const float cw = fmaxf(fmaxf(px1, tx1) - fminf(px0, tx0), 0.f), ch = fmaxf(fmaxf(py1, ty1) - fminf(py0, ty0), 0.f);
    const float carea = cw * ch;
    const float giou = iou - (carea - uni) / carea;
    const float cost_giou = -giou;
    const float x = a.logits[(int64_t)b * a.Q + q];
    float prob = sigmoidf_(x);
    float cost_class;
    if (!a.focal) {
      cost_class = -prob;
    } else if (a.stable) {
      prob = prob * ((-cost_giou + 1.f) / 2.f);
      cost_class = -a.alpha * powf(1.f - prob, a.gamma) * logf(prob) + (1.f - a.alpha) * powf(prob, a.gamma) * logf(1.f - prob);
    } else {
      cost_class = -a.alpha * powf(1.f - prob, a.gamma) * logsigmoidf_(x) + (1.f - a.alpha) * powf(prob, a.gamma) * logsigmoidf_(-x);
    }
    float c = (a.w_bbox * cost_bbox + a.w_class * cost_class) + a.w_giou * cost_giou;
    if (a.out_valid != nullptr && !a.out_valid[(int64_t)b * a.Q + q]) c = 1e9f;
    if (a.tgt_valid != nullptr && !a.tgt_valid[(int64_t)b * a.Tmax + t]) c = 1e9f;
    C[i] = c;
  }
}

struct Best { double val; int j; int is_new; };
// scipy: a column with a strictly lower path cost wins; at equal cost a column that is a new sink (unassigned) wins
__device__ __forceinline__ bool better(const Best& x, const Best& y) {
  if (x.j < 0) return false;
  if (y.j < 0) return true;
  if (x.val != y.val) return x.val < y.val;
  if (x.is_new != y.is_new) return x.is_new > y.is_new;
  return x.is_new ? (x.j < y.j) : (x.j > y.j);
}

// One CTA per image.  rows = the smaller side.  cost(i, j): rows are targets*repeats (transposed) when Tn <= Q, else queries.
__global__ void __launch_bounds__(256) matcher_lsa_kernel(const float* __restrict__ C, const int32_t* __restrict__ num_boxes, int Q,
                                                          int Tmax, int repeats, int filter, int32_t* __restrict__ query_of_col,
                                                          int32_t* __restrict__ col_of_query, int cmax) {
  extern __shared__ double shd[];
  const int b = blockIdx.x;
  const int T = min(num_boxes[b], Tmax);
  const int Tn = T * repeats;                        // columns of the reference's cost matrix after np.tile
  const float* Cb = C + (int64_t)b * Q * Tmax;
  int32_t* qoc = query_of_col + (int64_t)b * cmax;
  int32_t* coq = col_of_query + (int64_t)b * Q;
  for (int i = threadIdx.x; i < cmax; i += blockDim.x) qoc[i] = -1;
  for (int i = threadIdx.x; i < Q; i += blockDim.x) coq[i] = -1;
  if (Tn == 0) return;
  const bool tr = Tn < Q;                            // scipy transposes when nr > nc (nr == nc: rows stay the queries)
  const int R = tr ? Tn : Q, Cn = tr ? Q : Tn;
  auto cost = [&](int i, int j) -> double {
    const int q = tr ? j : i, c = tr ? i : j;
    return (double)Cb[(int64_t)q * Tmax + (c % T)];
  };
  const int maxn = max(Q, cmax);
  double* u = shd;                 // [R]
  double* v = u + maxn;            // [Cn]
  double* spc = v + maxn;          // [Cn] shortest path costs
  int* path = reinterpret_cast<int*>(spc + maxn);   // [Cn]
  int* row4col = path + maxn;      // [Cn]
  int* col4row = row4col + maxn;   // [R]
  unsigned char* SR = reinterpret_cast<unsigned char*>(col4row + maxn);   // [R]
  unsigned char* SC = SR + maxn;   // [Cn]
  __shared__ Best wbest[8];
  __shared__ int s_i, s_sink;
  __shared__ double s_min;
  for (int k = threadIdx.x; k < maxn; k += blockDim.x) { u[k] = 0.0; v[k] = 0.0; row4col[k] = -1; col4row[k] = -1; }
  __syncthreads();
  const double INF = DBL_MAX;
  for (int cur = 0; cur < R; ++cur) {
    for (int k = threadIdx.x; k < maxn; k += blockDim.x) { spc[k] = INF; SR[k] = 0; SC[k] = 0; path[k] = -1; }
    if (threadIdx.x == 0) { s_i = cur; s_sink = -1; s_min = 0.0; }
    __syncthreads();
    while (true) {
      const int i = s_i;
      const double minVal = s_min, ui = u[i];
      Best mine{INF, -1, 0};
      for (int j = threadIdx.x; j < Cn; j += blockDim.x) {
        if (SC[j]) continue;
        const double r = minVal + cost(i, j) - ui - v[j];
        if (r < spc[j]) { spc[j] = r; path[j] = i; }
        const Best cand{spc[j], j, row4col[j] == -1 ? 1 : 0};
        if (better(cand, mine)) mine = cand;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Best other;
        other.val = __shfl_xor_sync(0xffffffffu, mine.val, o);
        other.j = __shfl_xor_sync(0xffffffffu, mine.j, o);
        other.is_new = __shfl_xor_sync(0xffffffffu, mine.is_new, o);
        if (better(other, mine)) mine = other;
      }
      if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = mine;
      __syncthreads();
      if (threadIdx.x == 0) {
        Best bst = wbest[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
          if (better(wbest[w], bst)) bst = wbest[w];
        SR[i] = 1;
        s_min = bst.val;
        const int j = bst.j;             // j < 0 only if every remaining cost is NaN (Cn >= R: a free column always exists)
        if (j < 0) {
          s_sink = -2;                   // non-finite costs (diverged model): leave this image unmatched instead of faulting
        } else {
          SC[j] = 1;
          if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
        }
      }
      __syncthreads();
      if (s_sink != -1) break;
    }
    if (s_sink < 0) break;               // uniform across the block (shared flag)
    const double minVal = s_min;
    // dual updates (rectangular_lsap.cpp: u[curRow] += minVal; u[i] += minVal - spc[col4row[i]] for i in SR; v[j] -= minVal - spc[j] for j in SC)
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
      if (i == cur) u[i] += minVal;
      else if (SR[i]) u[i] += minVal - spc[col4row[i]];
    }
    for (int j = threadIdx.x; j < Cn; j += blockDim.x)
      if (SC[j]) v[j] -= minVal - spc[j];
    __syncthreads();
    if (threadIdx.x == 0) {   // augment along the path
      int j = s_sink;
      while (true) {
        const int i = path[j];
        row4col[j] = i;
        const int prev = col4row[i];
        col4row[i] = j;
        j = prev;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }
  // write the matching in (query, column) terms; drop pairs through invalid (1e9) entries like _do_matching's filter
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const int j = col4row[i];
    if (j < 0) continue;
    const int q = tr ? j : i, c = tr ? i : j;
    if (filter && !(Cb[(int64_t)q * Tmax + (c % T)] < 1e8f)) continue;
    qoc[c] = q;
    coq[q] = c;
  }
}

}  // namespace

static int64_t matcher_lsa_smem(int Q, int cmax) {
  const int64_t maxn = std::max(Q, cmax);
  return maxn * (3 * 8 + 3 * 4 + 2) + 64;
}

int matcher_run(const MatcherArgs& a, float* cost, int32_t* query_of_col, int32_t* col_of_query, cudaStream_t s) {
  SAM3B_REQUIRE(a.B >= 0 && a.Q > 0 && a.Tmax >= 0 && a.repeats >= 1, "matcher: bad sizes B=%d Q=%d Tmax=%d repeats=%d", a.B, a.Q, a.Tmax, a.repeats);
  if (a.B == 0) return 0;
  SAM3B_REQUIRE(a.logits && a.pred_boxes && a.num_boxes && query_of_col && col_of_query, "matcher: null tensor");
  const int cmax = std::max(1, a.Tmax * a.repeats);
  if (a.Tmax > 0) {
    SAM3B_REQUIRE(a.tgt_boxes && cost, "matcher: null targets / cost buffer");
    const int64_t total = (int64_t)a.B * a.Q * a.Tmax;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 8);
    matcher_cost_kernel<<<blocks, 256, 0, s>>>(a, cost);
    SAM3B_LAUNCHED();
  }
  const int64_t smem = matcher_lsa_smem(a.Q, cmax);
  SAM3B_REQUIRE(smem <= 200 * 1024, "matcher: %d queries x %d target columns exceed the shared-memory working set", a.Q, cmax);
  static bool attr_set = false;
  if (!attr_set) {
    SAM3B_CHECK_CUDA(cudaFuncSetAttribute(matcher_lsa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int filter = (a.out_valid != nullptr || a.tgt_valid != nullptr) ? 1 : 0;
  matcher_lsa_kernel<<<a.B, 256, smem, s>>>(cost, a.num_boxes, a.Q, a.Tmax, a.repeats, filter, query_of_col, col_of_query, cmax);
  SAM3B_LAUNCHED();
  return 0;
}

}  // namespace sam3b
