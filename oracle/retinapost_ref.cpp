// retinapost_ref.cpp — CPU ORACLE for the RetinaNet detection post-processing path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under oracle/ is part of the product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library, and only as
// the checker / the CPU baseline.  The product (retinanet-tensorflow2.x_b200/) never links, imports or calls it.
//
// *** PARITY UNPINNED (TF kernels) / PINNED (Python glue). ***  The reference
// (srihari-humbarwadi/retinanet-tensorflow2.x) holds no tests or golden vectors for this path, and its
// arithmetic lives in TensorFlow (tf-nightly 2.8.0-dev20210925: NonMaxSuppressionV5, CombinedNonMaxSuppression,
// TopKV2, Eigen sigmoid/exp), which is neither vendored in /root/reference nor installable here.  The TF kernels
// are therefore RESTATED from their published algorithm and checked against hand-derived
// known-answer tests, torchvision's NMS and the expectations of TensorFlow's own unit tests
// for these kernels (restated from the upstream test files: tests/test_tf_published_vectors.py) — nothing more.  The reference's own Python glue
// (mode dispatch, clipping, padding, dtypes, gather/top-k composition) IS pinned: tests/golden/ holds outputs of
// the unmodified reference modules executed here over a numpy stand-in for the handful of tf.* ops they call
// (tests/golden/make_golden.py), and this oracle is checked against them.
//
// Each function cites the reference file:line (relative to /root/reference) it follows.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -pthread (see oracle/Makefile).  -ffp-contract=off
// matters: the TF kernels are compiled without FMA contraction across statements and the IoU / decode
// arithmetic must round after every operation.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <numeric>
#include <queue>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------

template <class F>
void parallel_for(long n, int threads, F fn) {
  if (threads <= 1 || n <= 1) {
    for (long i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<long> next(0);
  std::vector<std::thread> pool;
  int t = (int)std::min<long>(threads, n);
  for (int w = 0; w < t; ++w)
    pool.emplace_back([&]() {
      for (;;) {
        long i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto& th : pool) th.join();
}

// tf.nn.sigmoid (postprocessing_ops.py:114).  TF-CPU evaluates Eigen's fp32 rational approximation (a few ulp
// from the exact value); the oracle states the function itself: the exact logistic evaluated in
// binary64 and rounded once to binary32.  Stage-1 parity is a tolerance (1e-5 relative), not bit-exactness.
inline float sigmoid_ref(float x) { return (float)(1.0 / (1.0 + std::exp(-(double)x))); }

// tf.math.exp (postprocessing_ops.py:97): same convention — exact exp rounded once to binary32.
inline float exp_ref(float x) { return (float)std::exp((double)x); }

inline float clip01(float v) { return std::min(std::max(v, 0.0f), 1.0f); }  // tf.clip_by_value(x, 0, 1)

// IoU of TF's NMS kernels (non_max_suppression_op.cc IOU).  All fp32, this association.
inline float iou_ref(const float* a, const float* b) {
  const float ymin_i = std::min(a[0], a[2]), xmin_i = std::min(a[1], a[3]);
  const float ymax_i = std::max(a[0], a[2]), xmax_i = std::max(a[1], a[3]);
  const float ymin_j = std::min(b[0], b[2]), xmin_j = std::min(b[1], b[3]);
  const float ymax_j = std::max(b[0], b[2]), xmax_j = std::max(b[1], b[3]);
  const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
  const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
  if (area_i <= 0 || area_j <= 0) return 0.0f;
  const float iymin = std::max(ymin_i, ymin_j), ixmin = std::max(xmin_i, xmin_j);
  const float iymax = std::min(ymax_i, ymax_j), ixmax = std::min(xmax_i, xmax_j);
  const float inter = std::max(iymax - iymin, 0.0f) * std::max(ixmax - ixmin, 0.0f);
  return inter / (area_i + area_j - inter);
}

// ------------------------------------------------------------------------------------------------------------
// TopKV2 CPU (topk_op.cc + lib/gtl/top_n.h) — called at postprocessing_ops.py:135,155,350,475.
// "better" = higher value, ties -> lower index.  sorted=true: best first.  sorted=false: gtl::TopN's internal
// array order (libstdc++ make_heap/pop_heap replayed).  k == n: full sort regardless of `sorted`.
// ------------------------------------------------------------------------------------------------------------
void topk_row(const float* v, int n, int k, bool sorted, int* idx_out) {
  auto better = [v](int a, int b) { return v[a] > v[b] || (v[a] == v[b] && a < b); };
  if (k <= 0) return;
  if (k >= n) {
    std::iota(idx_out, idx_out + n, 0);
    std::sort(idx_out, idx_out + n, better);
    return;
  }
  if (sorted) {
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::partial_sort(idx.begin(), idx.begin() + k, idx.end(), better);
    std::copy(idx.begin(), idx.begin() + k, idx_out);
    return;
  }
  // gtl::TopN<int32, stable_comp>(k): elements_ grows to k+1, then make_heap (min-heap on "better": front is the
  // worst), pop_heap parks the worst in the spare last slot; each later push that beats front replaces it.
  std::vector<int> el;
  el.reserve(k + 1);
  bool heap = false;
  for (int c = 0; c < n; ++c) {
    if (!heap) {
      el.push_back(c);
      if ((int)el.size() == k + 1) {
        std::make_heap(el.begin(), el.end(), better);
        std::pop_heap(el.begin(), el.end(), better);
        heap = true;
      }
    } else if (better(c, el.front())) {
      el.back() = c;
      std::pop_heap(el.begin(), el.end(), better);
    }
  }
  if (heap) el.pop_back();
  std::copy(el.begin(), el.end(), idx_out);
}

// ------------------------------------------------------------------------------------------------------------
// NonMaxSuppressionV5 (non_max_suppression_op.cc DoNonMaxSuppressionOp) — called at
// postprocessing_ops.py:249 and :444 with pad_to_max_output_size=True.
// `soft_ignores_iou_threshold` = 1: TF >= 2.3 form (weight applies when is_soft || sim <= thr).
//                               = 0: pre-2.3 form (weight = sim <= thr ? exp(..) : 0).
// ------------------------------------------------------------------------------------------------------------
struct Cand {
  int box_index;
  float score;
  int suppress_begin_index;
};

int nms_v5(const float* boxes /*[n,4]*/, long box_stride, const float* scores, long score_stride, int n, int M,
           float iou_threshold, float score_threshold, float soft_nms_sigma, int soft_ignores_iou_threshold,
           int* sel_idx /*[M] padded 0*/, float* sel_scores /*[M] padded 0*/) {
  auto cmp = [](const Cand& a, const Cand& b) {
    return ((a.score == b.score) && (a.box_index > b.box_index)) || a.score < b.score;
  };
  std::priority_queue<Cand, std::deque<Cand>, decltype(cmp)> pq(cmp);
  for (int i = 0; i < n; ++i) {
    float s = scores[(long)i * score_stride];
    if (s > score_threshold) pq.push(Cand{i, s, 0});
  }
  float scale = 0.0f;
  const bool is_soft = soft_nms_sigma > 0.0f;
  if (is_soft) scale = -0.5f / soft_nms_sigma;
  auto weight = [&](float sim) -> float {
    const float w = expf(scale * sim * sim);  // Eigen::numext::exp<float> -> libm expf
    if (soft_ignores_iou_threshold) return (is_soft || sim <= iou_threshold) ? w : 0.0f;
    return (sim <= iou_threshold) ? w : 0.0f;
  };
  std::vector<int> selected;
  std::vector<float> selected_scores;
  while ((int)selected.size() < M && !pq.empty()) {
    Cand c = pq.top();
    const float original = c.score;
    pq.pop();
    bool hard = false;
    for (int j = (int)selected.size() - 1; j >= c.suppress_begin_index; --j) {
      const float sim = iou_ref(boxes + (long)c.box_index * box_stride, boxes + (long)selected[j] * box_stride);
      c.score *= weight(sim);
      if (!is_soft && sim > iou_threshold) {
        hard = true;
        break;
      }
      if (c.score <= score_threshold) break;
    }
    c.suppress_begin_index = (int)selected.size();
    if (!hard) {
      if (c.score == original) {
        selected.push_back(c.box_index);
        selected_scores.push_back(c.score);
        continue;
      }
      if (c.score > score_threshold) pq.push(c);
    }
  }
  const int valid = (int)selected.size();
  for (int i = 0; i < M; ++i) {
    sel_idx[i] = i < valid ? selected[i] : 0;
    sel_scores[i] = i < valid ? selected_scores[i] : 0.0f;
  }
  return valid;
}

// ------------------------------------------------------------------------------------------------------------
// tf.image.non_max_suppression_padded (tensorflow/python/ops/image_ops_impl.py: non_max_suppression_padded ->
// non_max_suppression_padded_v2 with its helpers _bbox_overlap, _self_suppression, _cross_suppression,
// _suppression_loop_body) for ONE image, as the TPU branches call it (postprocessing_ops.py:323-330,
// :392-400): canonicalized_coordinates=True (no min/max flip), sorted_input=False, pad_to_max_output_size=True,
// tile_size=512.  TensorFlow is not in /root/reference: restated from the published algorithm, tile structure
// included, so that the CUDA path (a plain greedy scan) is checked against the reference's formulation:
//   1. optional score filter: scores *= (score > thr), boxes *= (score > thr);
//   2. argsort descending (top_k: ties -> lower index);
//   3. pad to a multiple of 512 boxes with zeros; while output_size < M and tiles remain:
//        cross-suppress the tile against every earlier tile (kept boxes only, the others are already zeroed), then
//        iterate the in-tile self-suppression to its fixed point, zero the suppressed boxes, count the boxes with
//        any coordinate > 0;
//   4. the first M non-zero boxes are the selection; indices beyond num_valid = min(output_size, M) are 0.
// IoU (_bbox_overlap): inter / (area_a + area_b - inter + 1e-8), fp32; a box suppresses when iou >= threshold
// (the NonMaxSuppressionV5 kernel uses a strict >).
// exact_fixed_point = 1: the self-suppression loop runs until no row changes.  0: TF's own stop test
// `iou_sum - iou_sum_new > iou_threshold` on fp32 sums (it can stop one round early when the removed IoU mass is
// within rounding of the threshold; the sums' association is the framework's, here row-major sequential).
// Returns num_valid; idx_out [M].  The batched loop condition of TF (`reduce_min(output_size) < M`) only makes
// finished images idle along: per-image results are independent.
// ------------------------------------------------------------------------------------------------------------
inline float iou_padded_ref(const float* a, const float* b) {
  const float i_xmin = std::max(a[1], b[1]), i_xmax = std::min(a[3], b[3]);
  const float i_ymin = std::max(a[0], b[0]), i_ymax = std::min(a[2], b[2]);
  const float i_area = std::max(i_xmax - i_xmin, 0.0f) * std::max(i_ymax - i_ymin, 0.0f);
  const float a_area = (a[2] - a[0]) * (a[3] - a[1]);
  const float b_area = (b[2] - b[0]) * (b[3] - b[1]);
  const float u_area = a_area + b_area - i_area + 1e-8f;
  return i_area / u_area;
}

int nms_padded(const float* boxes_in /*[n,4]*/, long box_stride, const float* scores_in, long score_stride, int n, int M,
               float iou_threshold, bool use_score_threshold, float score_threshold, int exact_fixed_point,
               int* idx_out /*[M]*/) {
  const int TS = 512;
  std::vector<float> sc(n);
  std::vector<float> bx((size_t)n * 4);
  for (int i = 0; i < n; ++i) {
    float s = scores_in[(long)i * score_stride];
    const float* b = boxes_in + (long)i * box_stride;
    float m = 1.0f;
    if (use_score_threshold) m = s > score_threshold ? 1.0f : 0.0f;   // filter_by_score
    sc[i] = use_score_threshold ? s * m : s;
    for (int k = 0; k < 4; ++k) bx[(size_t)i * 4 + k] = use_score_threshold ? b[k] * m : b[k];
  }
  std::vector<int> order(n);   // _sort_scores_and_boxes: argsort DESCENDING = top_k(k = n)
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int d) { return sc[a] > sc[d]; });
  const int nbp = (std::max(n, M) + TS - 1) / TS * TS;   // num_boxes_after_padding
  std::vector<float> sb((size_t)nbp * 4, 0.0f);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) sb[(size_t)i * 4 + k] = bx[(size_t)order[i] * 4 + k];
  const int num_iterations = nbp / TS;
  int output_size = 0;
  std::vector<float> iou((size_t)TS * TS);
  for (int idx = 0; idx < num_iterations && output_size < M; ++idx) {   // _suppression_loop_body
    float* slice = sb.data() + (size_t)idx * TS * 4;
    for (int inner = 0; inner < idx; ++inner) {                         // _cross_suppression
      const float* prev = sb.data() + (size_t)inner * TS * 4;
      for (int j = 0; j < TS; ++j) {
        bool all_below = true;
        for (int i = 0; i < TS && all_below; ++i)
          if (!(iou_padded_ref(prev + (size_t)i * 4, slice + (size_t)j * 4) < iou_threshold)) all_below = false;
        if (!all_below) slice[j * 4] = slice[j * 4 + 1] = slice[j * 4 + 2] = slice[j * 4 + 3] = 0.0f;
      }
      // (TF multiplies the slice by the 0/1 mask after each earlier tile; zeroing at once is the same because a
      // zeroed box has IoU 0 with everything and iou_threshold > 0)
    }
    for (int i = 0; i < TS; ++i)
      for (int j = 0; j < TS; ++j) {
        float v = 0.0f;
        if (j > i) {
          v = iou_padded_ref(slice + (size_t)i * 4, slice + (size_t)j * 4);
          if (!(v >= iou_threshold)) v = 0.0f;
        }
        iou[(size_t)i * TS + j] = v;
      }
    auto total = [&]() { float t = 0.0f; for (float v : iou) t += v; return t; };
    float iou_sum = total();
    for (;;) {                                                          // _self_suppression
      std::vector<char> can_suppress_others(TS), keep_row(TS);
      for (int j = 0; j < TS; ++j) {
        float mx = 0.0f;
        for (int i = 0; i < TS; ++i) mx = std::max(mx, iou[(size_t)i * TS + j]);
        can_suppress_others[j] = mx < iou_threshold;
      }
      for (int j = 0; j < TS; ++j) {
        float mx = 0.0f;
        for (int i = 0; i < TS; ++i)
          if (can_suppress_others[i]) mx = std::max(mx, iou[(size_t)i * TS + j]);
        keep_row[j] = mx < iou_threshold;
      }
      bool changed = false;
      for (int j = 0; j < TS; ++j)
        if (!keep_row[j])
          for (int k = 0; k < TS; ++k)
            if (iou[(size_t)j * TS + k] != 0.0f) { iou[(size_t)j * TS + k] = 0.0f; changed = true; }
      const float iou_sum_new = total();
      const bool again = exact_fixed_point ? changed : (iou_sum - iou_sum_new > iou_threshold);
      iou_sum = iou_sum_new;
      if (!again) break;
    }
    for (int j = 0; j < TS; ++j) {
      float col = 0.0f;
      for (int i = 0; i < TS; ++i) col += iou[(size_t)i * TS + j];
      if (col > 0.0f) slice[j * 4] = slice[j * 4 + 1] = slice[j * 4 + 2] = slice[j * 4 + 3] = 0.0f;   // suppressed_box
    }
    for (int j = 0; j < TS; ++j)
      if (slice[j * 4] > 0.0f || slice[j * 4 + 1] > 0.0f || slice[j * 4 + 2] > 0.0f || slice[j * 4 + 3] > 0.0f)
        ++output_size;
  }
  const int num_valid = std::min(output_size, M);
  // idx = nbp - top_k(any(selected > 0) * range(nbp, 0, -1), M): positions of the first M non-zero boxes
  int found = 0;
  for (int i = 0; i < nbp && found < M; ++i) {
    const float* b = sb.data() + (size_t)i * 4;
    if (b[0] > 0.0f || b[1] > 0.0f || b[2] > 0.0f || b[3] > 0.0f) {
      const int pos = std::min(i, n - 1);
      idx_out[found] = found < num_valid ? order[pos] : 0;
      ++found;
    }
  }
  for (int i = found; i < M; ++i) idx_out[i] = 0;
  for (int i = num_valid; i < M; ++i) idx_out[i] = 0;
  return num_valid;
}

struct Det {
  float score;
  int cls;
  int box_index;
  float box[4];
};

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------------------
// a1  AnchorBoxGenerator (dataloader/anchor_generator.py:24-104).  Returns N; writes [N,4] = [cx,cy,w,h].
// boundaries_out (optional) gets num_levels+1 cumulative counts (anchor_generator.py:42-49).
// ------------------------------------------------------------------------------------------------------------
long rpp_ref_anchors(int H, int W, int min_level, int max_level, const double* areas, int n_areas,
                     const double* ratios, int n_ratios, const double* scales, int n_scales, float* out,
                     long* boundaries_out) {
  const int A = n_ratios * n_scales;
  long n = 0;
  if (boundaries_out) boundaries_out[0] = 0;
  for (int l = min_level; l <= max_level; ++l) {
    const int li = l - min_level;
    const double stride_d = std::pow(2.0, l);
    const int fh = (int)std::ceil(H / stride_d), fw = (int)std::ceil(W / stride_d);  // :97-100
    if (out) {
      if (li >= n_areas) return -1;
      // _compute_dims :51-63 — area/ratio in Python float64, cast to fp32, sqrt/div/mul in fp32; ratio-major,
      // scale-minor.
      std::vector<float> dims(2 * A);
      int a = 0;
      for (int r = 0; r < n_ratios; ++r) {
        const float h = std::sqrt((float)(areas[li] / ratios[r]));
        const float w = (float)areas[li] / h;
        for (int s = 0; s < n_scales; ++s, ++a) {
          dims[2 * a + 0] = (float)scales[s] * w;
          dims[2 * a + 1] = (float)scales[s] * h;
        }
      }
      const float stride = (float)stride_d;
      for (int y = 0; y < fh; ++y)
        for (int x = 0; x < fw; ++x)
          for (int k = 0; k < A; ++k) {  // _get_anchors :78-88 — meshgrid(rx, ry): x fastest, then anchors
            float* o = out + (n + ((long)y * fw + x) * A + k) * 4;
            o[0] = ((float)x + 0.5f) * stride;
            o[1] = ((float)y + 0.5f) * stride;
            o[2] = dims[2 * k + 0];
            o[3] = dims[2 * k + 1];
          }
    }
    n += (long)fh * fw * A;
    if (boundaries_out) boundaries_out[li + 1] = n;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------------------
// a3  TransformBoxesAndScores (postprocessing_ops.py:87-117)
// ------------------------------------------------------------------------------------------------------------
void rpp_ref_sigmoid(const float* x, float* y, long n, int threads) {
  const long chunk = 1 << 16;
  parallel_for((n + chunk - 1) / chunk, threads, [&](long t) {
    const long e = std::min(n, (t + 1) * chunk);
    for (long i = t * chunk; i < e; ++i) y[i] = sigmoid_ref(x[i]);
  });
}

void rpp_ref_decode_boxes(const float* deltas /*[B,N,4]*/, const float* anchors /*[N,4] cx,cy,w,h*/, long B, long N,
                          int H, int W, const float* box_variance /*[4]*/, int scale_box_targets,
                          float* boxes_out /*[B,N,4]*/, int threads) {
  const float shape[4] = {(float)H, (float)W, (float)H, (float)W};  // :65-69 tile([H,W],2) applied to x1,y1,x2,y2
  parallel_for(B, threads, [&](long b) {
    for (long i = 0; i < N; ++i) {
      float d[4];
      for (int k = 0; k < 4; ++k) {
        d[k] = deltas[(b * N + i) * 4 + k];
        if (scale_box_targets) d[k] = d[k] * box_variance[k];  // :90-91
      }
      const float* a = anchors + i * 4;
      float xy[2], half[2];
      for (int k = 0; k < 2; ++k) {
        const float m = d[k] * a[2 + k];        // :96 boxes_xy * anchors_wh
        xy[k] = m + a[k];                       //     + anchors_xy
        const float wh = exp_ref(d[2 + k]) * a[2 + k];  // :97
        half[k] = wh / 2.0f;                    // :98
      }
      float* o = boxes_out + (b * N + i) * 4;
      o[0] = (xy[0] - half[0]) / shape[0];  // :100-104
      o[1] = (xy[1] - half[1]) / shape[1];
      o[2] = (xy[0] + half[0]) / shape[2];
      o[3] = (xy[1] + half[1]) / shape[3];
    }
  });
}

// ------------------------------------------------------------------------------------------------------------
// TopKV2 on a [rows, cols] matrix (exposed for tests).
// ------------------------------------------------------------------------------------------------------------
void rpp_ref_topk(const float* v, long rows, int cols, int k, int sorted, int* idx_out /*[rows,k']*/, int threads) {
  const int kk = std::min(k, cols);
  parallel_for(rows, threads, [&](long r) { topk_row(v + r * cols, cols, kk, sorted != 0, idx_out + r * kk); });
}

// ------------------------------------------------------------------------------------------------------------
// a4  FilterTopKDetections._filter_per_class (postprocessing_ops.py:128-147)
//     scores [B,N,C], boxes [B,N,4] -> scores_out [B,k',C], boxes_out [B,k',C,4], k' = min(k,N)
//     idx_out (optional) [B,C,k'] anchor indices.
// ------------------------------------------------------------------------------------------------------------
void rpp_ref_filter_per_class(const float* scores, const float* boxes, long B, long N, int C, int k, int sorted,
                              float* scores_out, float* boxes_out, int* idx_out, int threads) {
  const int kk = (int)std::min<long>(k, N);
  parallel_for(B * C, threads, [&](long t) {
    const long b = t / C;
    const int c = (int)(t % C);
    std::vector<float> col(N);  // :132-133 transpose + reshape to [B*C, N]
    for (long i = 0; i < N; ++i) col[i] = scores[(b * N + i) * C + c];
    std::vector<int> idx(kk);
    topk_row(col.data(), (int)N, kk, sorted != 0, idx.data());  // :135-138
    for (int j = 0; j < kk; ++j) {
      scores_out[(b * kk + j) * C + c] = col[idx[j]];  // :140-142
      const float* src = boxes + (b * N + idx[j]) * 4;  // :145 gather(boxes, indices, batch_dims=1)
      float* dst = boxes_out + ((b * kk + j) * C + c) * 4;
      dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
      if (idx_out) idx_out[(b * C + c) * kk + j] = idx[j];
    }
  });
}

// ------------------------------------------------------------------------------------------------------------
// a5  FilterTopKDetections._filter_global (postprocessing_ops.py:149-161)
//     -> scores_out [B,k',C], boxes_out [B,k',4], k' = min(k, N*C); idx_out (optional) [B,k'] flat indices.
// ------------------------------------------------------------------------------------------------------------
void rpp_ref_filter_global(const float* scores, const float* boxes, long B, long N, int C, int k, int sorted,
                           float* scores_out, float* boxes_out, int* idx_out, int threads) {
  const long NC = N * C;
  const int kk = (int)std::min<long>(k, NC);
  parallel_for(B, threads, [&](long b) {
    std::vector<int> idx(kk);
    topk_row(scores + b * NC, (int)NC, kk, sorted != 0, idx.data());  // :153-155
    for (int j = 0; j < kk; ++j) {
      const long a = idx[j] / C;  // :156
      std::memcpy(scores_out + (b * kk + j) * C, scores + (b * N + a) * C, sizeof(float) * C);  // :158
      std::memcpy(boxes_out + (b * kk + j) * 4, boxes + (b * N + a) * 4, sizeof(float) * 4);    // :159
      if (idx_out) idx_out[b * kk + j] = idx[j];
    }
  });
}

// ------------------------------------------------------------------------------------------------------------
// NonMaxSuppressionV5 on one problem (exposed for the KATs of SURVEY.md D.3).
// ------------------------------------------------------------------------------------------------------------
int rpp_ref_nms_v5(const float* boxes, const float* scores, int n, int M, float iou_threshold, float score_threshold,
                   float soft_nms_sigma, int soft_ignores_iou_threshold, int* sel_idx, float* sel_scores) {
  return nms_v5(boxes, 4, scores, 1, n, M, iou_threshold, score_threshold, soft_nms_sigma,
                soft_ignores_iou_threshold, sel_idx, sel_scores);
}

float rpp_ref_iou(const float* a, const float* b) { return iou_ref(a, b); }

// ------------------------------------------------------------------------------------------------------------
// a6-a9  GenerateDetections.call (postprocessing_ops.py:537-561) on non-TPU strategies.
//   mode: 0 CombinedNMS, 1 GlobalSoftNMS, 2 GlobalHardNMS, 3 PerClassSoftNMS, 4 PerClassHardNMS
//   scores [B,n,C]; boxes [B,n,q,4] (q = 1 for 3-D boxes, q = C after the per-class filter)
//   outputs: boxes_out [B,M,4] f32, scores_out [B,M] f32, classes_out [B,M] (f32 for mode 0, int64 for 1-2,
//   int32 for 3-4), valid_out [B] i32.
//   topk_sorted: 1 = canonical (sorted=True) order for the final tf.nn.top_k of the PerClass modes (:475-477);
//                0 = TF-CPU sorted=False heap order.
// Returns 0, or -1 for an invalid mode/rank combination (Global* with q != 1: rank error in TF).
// ------------------------------------------------------------------------------------------------------------
int rpp_ref_generate_detections(int mode, const float* scores, const float* boxes_in, long B, long n, int q, int C,
                                float iou_threshold, float score_threshold, int M, float sigma, int topk_sorted,
                                int soft_ignores_iou_threshold, float* boxes_out, float* scores_out,
                                void* classes_out, int* valid_out, int threads) {
  if (mode < 0 || mode > 4) return -1;
  if ((mode == 1 || mode == 2) && q != 1) return -1;

  if (mode == 0) {
    // _combined_nms :219-242 -> tf.image.combined_non_max_suppression(max_output_size_per_class=M,
    // max_total_size=M, clip_boxes=True, pad_per_class=False).  SURVEY.md A.3.  Ties canonicalised:
    // per-class queue (score desc, box index asc); cross-class merge (score desc, class asc, selection order).
    const int size_per_class = (int)std::min<long>(M, n);
    std::vector<std::vector<Det>> per_task(B * C);
    parallel_for(B * C, threads, [&](long t) {
      const long b = t / C;
      const int c = (int)(t % C);
      const int qi = q > 1 ? c : 0;
      std::vector<int> cand;
      for (long i = 0; i < n; ++i)
        if (scores[(b * n + i) * C + c] > score_threshold) cand.push_back((int)i);
      auto sc = [&](int i) { return scores[(b * n + i) * C + c]; };
      std::sort(cand.begin(), cand.end(), [&](int a, int d) { return sc(a) > sc(d) || (sc(a) == sc(d) && a < d); });
      std::vector<Det>& kept = per_task[t];
      for (int i : cand) {
        if ((int)kept.size() >= size_per_class) break;
        const float* bx = boxes_in + ((b * n + i) * q + qi) * 4;  // raw, unclipped (B6)
        bool ok = true;
        for (int j = (int)kept.size() - 1; j >= 0; --j)
          if (iou_ref(bx, kept[j].box) > iou_threshold) { ok = false; break; }
        if (ok) kept.push_back(Det{sc(i), c, i, {bx[0], bx[1], bx[2], bx[3]}});
      }
    });
    float* classes = (float*)classes_out;
    parallel_for(B, threads, [&](long b) {
      std::vector<Det> all;
      for (int c = 0; c < C; ++c) all.insert(all.end(), per_task[b * C + c].begin(), per_task[b * C + c].end());
      std::stable_sort(all.begin(), all.end(), [](const Det& a, const Det& d) { return a.score > d.score; });
      const int valid = (int)std::min<size_t>(all.size(), M);
      valid_out[b] = valid;
      for (int i = 0; i < M; ++i) {
        float* bo = boxes_out + (b * M + i) * 4;
        if (i < valid) {
          for (int k = 0; k < 4; ++k) bo[k] = clip01(all[i].box[k]);
          scores_out[b * M + i] = all[i].score;
          classes[b * M + i] = (float)all[i].cls;
        } else {
          bo[0] = bo[1] = bo[2] = bo[3] = 0.0f;
          scores_out[b * M + i] = 0.0f;
          classes[b * M + i] = 0.0f;
        }
      }
    });
    return 0;
  }

  // every other mode clips the boxes first (:275, :501)
  std::vector<float> boxes((size_t)B * n * q * 4);
  {
    const long tot = B * n * q * 4, chunk = 1 << 16;
    parallel_for((tot + chunk - 1) / chunk, threads, [&](long t) {
      const long e = std::min(tot, (t + 1) * chunk);
      for (long i = t * chunk; i < e; ++i) boxes[i] = clip01(boxes_in[i]);
    });
  }

  if (mode == 1 || mode == 2) {
    // _global_nms :244-286.  GlobalHardNMS dispatches with sigma=0.0 (:552): `1.0 if not sigma` -> IoU threshold
    // 1.0 (B1); GlobalSoftNMS passes the real threshold (B2) and soft_nms_sigma = sigma/2 (B4).
    const float sg = (mode == 2) ? 0.0f : sigma;
    const float iou_thr = (sg == 0.0f) ? 1.0f : iou_threshold;
    long long* classes = (long long*)classes_out;
    parallel_for(B, threads, [&](long b) {
      std::vector<float> s(n);
      std::vector<int> cls(n);
      for (long i = 0; i < n; ++i) {  // reduce_max :251 / argmax :259 (first maximum)
        const float* row = scores + (b * n + i) * C;
        int best = 0;
        for (int c = 1; c < C; ++c)
          if (row[c] > row[best]) best = c;
        s[i] = row[best];
        cls[i] = best;
      }
      std::vector<int> sel(M);
      std::vector<float> sel_s(M);
      const int valid = nms_v5(boxes.data() + b * n * 4, 4, s.data(), 1, (int)n, M, iou_thr, score_threshold,
                               sg / 2.0f, soft_ignores_iou_threshold, sel.data(), sel_s.data());
      valid_out[b] = valid;
      for (int i = 0; i < M; ++i) {
        const float* bx = boxes.data() + (b * n + sel[i]) * 4;  // :258 gather (padded index 0 -> boxes[0], B8)
        float* bo = boxes_out + (b * M + i) * 4;
        bo[0] = bx[0]; bo[1] = bx[1]; bo[2] = bx[2]; bo[3] = bx[3];
        scores_out[b * M + i] = i < valid ? sel_s[i] : -1.0f;        // :262-264
        classes[b * M + i] = i < valid ? (long long)cls[sel[i]] : -1;  // :266-268
      }
    });
    return 0;
  }

  // _per_class_nms :434-535.  PerClassHardNMS dispatches sigma=0.0 (:561): `1.0 if sigma else iou` -> real
  // threshold; PerClassSoftNMS -> threshold 1.0 (B3), soft_nms_sigma = sigma/2.
  {
    const float sg = (mode == 4) ? 0.0f : sigma;
    const float iou_thr = (sg != 0.0f) ? 1.0f : iou_threshold;
    std::vector<float> f_scores((size_t)B * C * M), f_boxes((size_t)B * C * M * 4);
    parallel_for(B * C, threads, [&](long t) {
      const long b = t / C;
      const int c = (int)(t % C);
      const int qi = std::min(q - 1, c);  // :440
      std::vector<int> sel(M);
      const float* bx0 = boxes.data() + (b * n * q + qi) * 4;
      nms_v5(bx0, (long)q * 4, scores + b * n * C + c, C, (int)n, M, iou_thr, score_threshold, sg / 2.0f,
             soft_ignores_iou_threshold, sel.data(), f_scores.data() + t * M);
      for (int i = 0; i < M; ++i) {  // :453 gather (padded index 0 -> that class's box 0)
        const float* bx = bx0 + (long)sel[i] * q * 4;
        float* bo = f_boxes.data() + (t * M + i) * 4;
        bo[0] = bx[0]; bo[1] = bx[1]; bo[2] = bx[2]; bo[3] = bx[3];
      }
    });
    int* classes = (int*)classes_out;
    parallel_for(B, threads, [&](long b) {
      const int tot = C * M;
      std::vector<int> top(M);
      topk_row(f_scores.data() + b * tot, tot, M, topk_sorted != 0, top.data());  // :475-477
      int valid = 0;
      for (int i = 0; i < M; ++i)
        if (f_scores[b * tot + top[i]] > 0.0f) ++valid;  // :481-482
      valid_out[b] = valid;
      for (int i = 0; i < M; ++i) {
        const float* bx = f_boxes.data() + (b * tot + top[i]) * 4;  // :479
        float* bo = boxes_out + (b * M + i) * 4;
        bo[0] = bx[0]; bo[1] = bx[1]; bo[2] = bx[2]; bo[3] = bx[3];
        scores_out[b * M + i] = i < valid ? f_scores[b * tot + top[i]] : -1.0f;  // :484-486 (mask by POSITION, B9)
        classes[b * M + i] = i < valid ? top[i] / M : -1;                        // :468, :488-490
      }
    });
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// a10/a11  the TPUStrategy branches of GenerateDetections (postprocessing_ops.py:288-432), offered by the product as
// the opt-in `tpu_semantics` of the two hard modes.
//   mode 2 (GlobalHardNMS)   -> _tpu_global_hard_nms   :381-432: a real global hard NMS (true IoU threshold);
//   mode 4 (PerClassHardNMS) -> _tpu_per_class_hard_nms :288-379.
// Outputs: boxes [B,M,4], scores [B,M], classes [B,M] int32 (both), valid [B] i32; invalid positions are -1 in
// every field.
// ------------------------------------------------------------------------------------------------------------
int rpp_ref_nms_padded(const float* boxes, const float* scores, int n, int M, float iou_threshold,
                       int use_score_threshold, float score_threshold, int exact_fixed_point, int* idx_out) {
  return nms_padded(boxes, 4, scores, 1, n, M, iou_threshold, use_score_threshold != 0, score_threshold,
                    exact_fixed_point, idx_out);
}

float rpp_ref_iou_padded(const float* a, const float* b) { return iou_padded_ref(a, b); }

int rpp_ref_generate_detections_tpu(int mode, const float* scores, const float* boxes_in, long B, long n, int q, int C,
                                    float iou_threshold, float score_threshold, int M, int exact_fixed_point,
                                    float* boxes_out, float* scores_out, int* classes_out, int* valid_out,
                                    int threads) {
  if (mode != 2 && mode != 4) return -1;
  if (mode == 2 && q != 1) return -1;
  std::vector<float> boxes((size_t)B * n * q * 4);   // tf.clip_by_value :291, :385
  for (size_t i = 0; i < boxes.size(); ++i) boxes[i] = clip01(boxes_in[i]);

  if (mode == 2) {
    parallel_for(B, threads, [&](long b) {
      std::vector<float> s(n);
      std::vector<int> cls(n);
      for (long i = 0; i < n; ++i) {   // reduce_max :382 / argmax :383 (first maximum)
        const float* row = scores + (b * n + i) * C;
        int best = 0;
        for (int c = 1; c < C; ++c)
          if (row[c] > row[best]) best = c;
        s[i] = row[best];
        cls[i] = best;
      }
      std::vector<int> idx(M);
      const int valid = nms_padded(boxes.data() + b * n * 4, 4, s.data(), 1, (int)n, M, iou_threshold, true,
                                   score_threshold, exact_fixed_point, idx.data());   // :392-400
      valid_out[b] = valid;
      for (int i = 0; i < M; ++i) {    // :402-420: gather [boxes, classes, scores], -1 beyond valid
        float* bo = boxes_out + (b * M + i) * 4;
        if (i < valid) {
          const float* bx = boxes.data() + (b * n + idx[i]) * 4;
          bo[0] = bx[0]; bo[1] = bx[1]; bo[2] = bx[2]; bo[3] = bx[3];
          scores_out[b * M + i] = s[idx[i]];
          classes_out[b * M + i] = (int)(float)cls[idx[i]];   // cast f32 -> int32 (:425-426)
        } else {
          bo[0] = bo[1] = bo[2] = bo[3] = -1.0f;
          scores_out[b * M + i] = -1.0f;
          classes_out[b * M + i] = -1;
        }
      }
    });
    return 0;
  }

  std::vector<float> f_scores((size_t)B * C * M), f_boxes((size_t)B * C * M * 4);
  parallel_for(B * C, threads, [&](long t) {
    const long b = t / C;
    const int c = (int)(t % C);
    const int qi = q > 1 ? c : 0;    // boxes_idx :303-320
    std::vector<int> idx(M);
    const float* bx0 = boxes.data() + (b * n * q + qi) * 4;
    const float* sc0 = scores + b * n * C + c;
    nms_padded(bx0, (long)q * 4, sc0, C, (int)n, M, iou_threshold, false, 0.0f, exact_fixed_point,
               idx.data());           // :323-330: no score threshold inside
    for (int i = 0; i < M; ++i) {     // :332-335 gathers (padded index 0 -> that class's row 0, score included)
      const float* bx = bx0 + (long)idx[i] * q * 4;
      float* bo = f_boxes.data() + (t * M + i) * 4;
      bo[0] = bx[0]; bo[1] = bx[1]; bo[2] = bx[2]; bo[3] = bx[3];
      f_scores[t * M + i] = sc0[(long)idx[i] * C];
    }
  });
  parallel_for(B, threads, [&](long b) {
    const int tot = C * M;
    std::vector<int> top(M);
    topk_row(f_scores.data() + b * tot, tot, M, true, top.data());   // :350-351 tf.nn.top_k (sorted)
    int valid = 0;
    for (int i = 0; i < M; ++i) {
      const float sc = f_scores[b * tot + top[i]];
      const bool ok = sc > score_threshold;                           // :358-359
      float* bo = boxes_out + (b * M + i) * 4;
      if (ok) {
        const float* bx = f_boxes.data() + (b * tot + top[i]) * 4;
        bo[0] = bx[0]; bo[1] = bx[1]; bo[2] = bx[2]; bo[3] = bx[3];
        scores_out[b * M + i] = sc;
        classes_out[b * M + i] = top[i] / M;
        ++valid;
      } else {                                                        // :365-368
        bo[0] = bo[1] = bo[2] = bo[3] = -1.0f;
        scores_out[b * M + i] = -1.0f;
        classes_out[b * M + i] = -1;
      }
    }
    valid_out[b] = valid;                                             // :361-363
  });
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// a12  ModelBuilder.add_post_processing_stage (model/builder.py:153-190) after FuseDetections:
//      TransformBoxesAndScores -> [FilterTopKDetections if pre_nms_top_k > 0] -> GenerateDetections.
// class_logits [B,N,C], encoded_boxes [B,N,4], anchors [N,4].
// ------------------------------------------------------------------------------------------------------------
static int detect_common(const float* class_logits, const float* encoded_boxes, const float* anchors, long B, long N,
                         int C, int H, int W, const float* box_variance, int scale_box_targets, int mode,
                         float iou_threshold, float score_threshold, float sigma, int pre_nms_top_k,
                         int filter_per_class, int M, int topk_sorted, int soft_ignores_iou_threshold, int tpu,
                         float* boxes_out, float* scores_out, void* classes_out, int* valid_out, int threads) {
  auto gen = [&](const float* sc, const float* bx, long n, int q) -> int {
    if (tpu)
      return rpp_ref_generate_detections_tpu(mode, sc, bx, B, n, q, C, iou_threshold, score_threshold, M, 1,
                                             boxes_out, scores_out, (int*)classes_out, valid_out, threads);
    return rpp_ref_generate_detections(mode, sc, bx, B, n, q, C, iou_threshold, score_threshold, M, sigma, topk_sorted,
                                       soft_ignores_iou_threshold, boxes_out, scores_out, classes_out, valid_out,
                                       threads);
  };
  std::vector<float> scores((size_t)B * N * C), boxes((size_t)B * N * 4);
  rpp_ref_sigmoid(class_logits, scores.data(), B * N * C, threads);
  rpp_ref_decode_boxes(encoded_boxes, anchors, B, N, H, W, box_variance, scale_box_targets, boxes.data(), threads);
  if (pre_nms_top_k > 0) {
    if (filter_per_class) {
      const long kk = std::min<long>(pre_nms_top_k, N);
      std::vector<float> fs((size_t)B * kk * C), fb((size_t)B * kk * C * 4);
      rpp_ref_filter_per_class(scores.data(), boxes.data(), B, N, C, pre_nms_top_k, topk_sorted, fs.data(),
                               fb.data(), nullptr, threads);
      return gen(fs.data(), fb.data(), kk, C);
    }
    const long kk = std::min<long>(pre_nms_top_k, N * C);
    std::vector<float> fs((size_t)B * kk * C), fb((size_t)B * kk * 4);
    rpp_ref_filter_global(scores.data(), boxes.data(), B, N, C, pre_nms_top_k, topk_sorted, fs.data(), fb.data(),
                          nullptr, threads);
    return gen(fs.data(), fb.data(), kk, 1);
  }
  return gen(scores.data(), boxes.data(), N, 1);
}

int rpp_ref_detect(const float* class_logits, const float* encoded_boxes, const float* anchors, long B, long N, int C,
                   int H, int W, const float* box_variance, int scale_box_targets, int mode, float iou_threshold,
                   float score_threshold, float sigma, int pre_nms_top_k, int filter_per_class, int M,
                   int topk_sorted, int soft_ignores_iou_threshold, float* boxes_out, float* scores_out,
                   void* classes_out, int* valid_out, int threads) {
  return detect_common(class_logits, encoded_boxes, anchors, B, N, C, H, W, box_variance, scale_box_targets, mode,
                       iou_threshold, score_threshold, sigma, pre_nms_top_k, filter_per_class, M, topk_sorted,
                       soft_ignores_iou_threshold, 0, boxes_out, scores_out, classes_out, valid_out, threads);
}

// The same composition with the TPU branches of GenerateDetections (mode 2 or 4).
int rpp_ref_detect_tpu(const float* class_logits, const float* encoded_boxes, const float* anchors, long B, long N,
                       int C, int H, int W, const float* box_variance, int scale_box_targets, int mode,
                       float iou_threshold, float score_threshold, int pre_nms_top_k, int filter_per_class, int M,
                       float* boxes_out, float* scores_out, int* classes_out, int* valid_out, int threads) {
  if (mode != 2 && mode != 4) return -1;
  if (mode == 2 && pre_nms_top_k > 0 && filter_per_class) return -1;
  return detect_common(class_logits, encoded_boxes, anchors, B, N, C, H, W, box_variance, scale_box_targets, mode,
                       iou_threshold, score_threshold, 0.0f, pre_nms_top_k, filter_per_class, M, 1, 1, 1, boxes_out,
                       scores_out, classes_out, valid_out, threads);
}

// ------------------------------------------------------------------------------------------------------------
// §8f-4  EfficientNMS_TRT, the node the reference appends in export mode onnx_tensorrt (onnx_utils.py:13-85) with
// attributes score_activation=True, box_coding=1, background_class=-1, max_output_boxes=M.  The plugin is part of
// TensorRT (plugin/efficientNMSPlugin), which is neither in /root/reference nor installable here: PARITY UNPINNED.
// Restated from its published algorithm: filter sigmoid(logit) >= score_threshold, keep the 4096 best (anchor,
// class) pairs per image, decode centre-size boxes against the anchors (no variance, no normalisation), greedy
// class-aware NMS (IoU > threshold drops; IoU = inter / (a1 + a2 - inter), 0 when an area or the intersection is
// not positive), first M kept, outputs in centre-size coding, zero-filled beyond the count.  Ties are ordered by
// flat index (anchor * C + class); the plugin's own tie order is unspecified.
// ------------------------------------------------------------------------------------------------------------
int rpp_ref_efficient_nms(const float* raw_boxes /*[B,N,4]*/, const float* class_logits /*[B,N,C]*/,
                          const float* anchors /*[N,4] cx,cy,w,h*/, long B, long N, int C, int M,
                          float score_threshold, float iou_threshold, int* valid_out /*[B]*/,
                          float* boxes_out /*[B,M,4]*/, float* scores_out /*[B,M]*/, int* classes_out /*[B,M]*/,
                          int threads) {
  const long selected = 4096;
  parallel_for(B, threads, [&](long b) {
    std::vector<std::pair<float, long>> cand;
    for (long f = 0; f < N * C; ++f) {
      const float s = sigmoid_ref(class_logits[b * N * C + f]);
      if (s >= score_threshold) cand.emplace_back(s, f);
    }
    auto better = [](const std::pair<float, long>& a, const std::pair<float, long>& d) {
      return a.first > d.first || (a.first == d.first && a.second < d.second);
    };
    const long take = std::min<long>(selected, (long)cand.size());
    std::partial_sort(cand.begin(), cand.begin() + take, cand.end(), better);
    struct Kept { float box[4]; float area; int cls; };
    std::vector<Kept> kept;
    for (int i = 0; i < M; ++i) {
      float* bo = boxes_out + (b * M + i) * 4;
      bo[0] = bo[1] = bo[2] = bo[3] = 0.0f;
      scores_out[b * M + i] = 0.0f;
      classes_out[b * M + i] = 0;
    }
    for (long i = 0; i < take && (int)kept.size() < M; ++i) {
      const long row = cand[i].second / C;
      const int cls = (int)(cand[i].second % C);
      const float* d = raw_boxes + (b * N + row) * 4;
      const float* a = anchors + row * 4;
      const float cx = d[0] * a[2] + a[0], cy = d[1] * a[3] + a[1];
      const float hw = (a[2] * exp_ref(d[2])) * 0.5f, hh = (a[3] * exp_ref(d[3])) * 0.5f;
      Kept k{{cx - hw, cy - hh, cx + hw, cy + hh}, 0.0f, cls};
      const float w = k.box[2] - k.box[0], h = k.box[3] - k.box[1];
      k.area = (w > 0.0f && h > 0.0f) ? w * h : 0.0f;
      bool ok = true;
      for (const Kept& o : kept) {
        if (o.cls != cls || !(k.area > 0.0f) || !(o.area > 0.0f)) continue;
        const float iw = std::min(k.box[2], o.box[2]) - std::max(k.box[0], o.box[0]);
        const float ih = std::min(k.box[3], o.box[3]) - std::max(k.box[1], o.box[1]);
        if (!(iw > 0.0f && ih > 0.0f)) continue;
        const float inter = iw * ih;
        const float uni = k.area + o.area - inter;
        if (uni > 0.0f && inter / uni > iou_threshold) { ok = false; break; }
      }
      if (!ok) continue;
      const int pos = (int)kept.size();
      float* bo = boxes_out + (b * M + pos) * 4;
      bo[0] = k.box[0] + 0.5f * w; bo[1] = k.box[1] + 0.5f * h; bo[2] = w; bo[3] = h;
      scores_out[b * M + pos] = cand[i].first;
      classes_out[b * M + pos] = cls;
      kept.push_back(k);
    }
    valid_out[b] = (int)kept.size();
  });
  return 0;
}

int rpp_ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
