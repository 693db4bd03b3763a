// rpp_outputs.cuh — K4-K8: per-image merges, Global* outputs, top-k gathers, EfficientNMS entry, COCO epilogue.
// Part of the retinapost kernel set; included by rpp_kernels.cuh (one translation unit: rpp_api.cu).
#pragma once
#include "rpp_kernels.cuh"

// ===============================================================================================================
// K4  per-image merge (PerClass*: concat C*M + tf.nn.top_k(M) + positional mask, postprocessing_ops.py:471-490;
//     CombinedNMS: SelectResultPerBatch).  One block per image.
// ===============================================================================================================
#define RPP_MERGE_NT 256

struct MergeParams {
  int C, M;
  int combined;            // 1: CombinedNMS output convention, 0: PerClass*
  const u64* sel_key;      // [B*C][M]
  const float4* sel_box;   // [B*C][M]
  const int* sel_cnt;      // [B*C]
  // pad box of a class with no candidates = its row 0 (:453 gather of index 0): needs the column argmax
  Levels lv; int is_logit; long N;
  const float4* anchors; const float4* boxes; int q; DecodeParams dp;
  int row0_mode;           // 0: row 0 = index 0 of the source; 1: row 0 = best of the column (per-class top-k ran)
  int keys_in_smem;        // the C*M merge keys fit in dynamic shared memory
  int score_nonneg;        // score_threshold >= 0: every kept score is positive
  float4* out_boxes;       // [B][M]
  float* out_scores;       // [B][M]
  void* out_classes;       // [B][M] f32 (combined) / i32
  int* out_valid;          // [B]
};

struct MergeShared {
  SelectScratch<RPP_MERGE_NT> sel;
  u64 chunk[RPP_CHUNK_CAP];
  u64 top[1024];
  int need[1024];
  int npos;
};

__global__ void __launch_bounds__(RPP_MERGE_NT) merge_kernel(MergeParams P) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MergeShared* sh = reinterpret_cast<MergeShared*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int C = P.C, M = P.M;
  const u64* sk = P.sel_key + (size_t)b * C * M;
  const int* scnt = P.sel_cnt + (size_t)b * C;
  auto keyfn = [&](int i) -> u64 {
    const int c = i / M, slot = i - c * M;
    if (slot < scnt[c]) return make_key(key_score(sk[i]), (u32)i);
    return P.combined ? 0ull : make_key(0.0f, (u32)i);  // NMSV5 pads scores with 0.0 (A.2)
  };
  u64* skeys = reinterpret_cast<u64*>(sh + 1);
  int got = 0;
  // Fast path (the usual case after the cross-class bound: a few boxes per class, at least M in total and no more than
  // the chunk buffer holds): compact the kept boxes' keys and sort them once.
  bool fast = false;
  {
    // per-class counts -> exclusive offsets (thread 0; C is at most a few thousand)
    __shared__ int s_total;
    int* s_pref = reinterpret_cast<int*>(sh->top);   // top[] (1024 u64 = 2048 ints) is not live yet
    const bool fits = C <= 2047;
    if (fits) {
      for (int c = tid; c < C; c += RPP_MERGE_NT) s_pref[c + 1] = scnt[c] < M ? scnt[c] : M;
      __syncthreads();
      if (tid == 0) {
        int run = 0;
        for (int c = 0; c < C; ++c) { const int n = s_pref[c + 1]; s_pref[c] = run; run += n; }
        s_pref[C] = run;
        s_total = run;
      }
      __syncthreads();
      const int total = s_total;
      // (PerClass*: the zero-score pads of NMSV5 can only matter when a kept score may be <= 0, i.e. with a
      // negative score threshold; those cases take the general path)
      fast = total >= M && total <= RPP_CHUNK_CAP && (P.combined || P.score_nonneg);
      if (fast) {
        for (int c = tid; c < C; c += RPP_MERGE_NT) {
          const int o = s_pref[c], n = s_pref[c + 1] - o;
          for (int slot = 0; slot < n; ++slot)
            sh->chunk[o + slot] = make_key(key_score(sk[(size_t)c * M + slot]), (u32)(c * M + slot));
        }
        const int P2 = next_pow2(total < 2 ? 2 : total);
        for (int i = total + tid; i < P2; i += RPP_MERGE_NT) sh->chunk[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc<RPP_MERGE_NT>(sh->chunk, P2);
        for (int i = tid; i < M; i += RPP_MERGE_NT) sh->top[i] = sh->chunk[i];   // s_pref is dead from here on
        got = M;
        __syncthreads();
      }
    }
  }
  // General path: stage the C*M keys (pads included) in shared memory once, then select over them.
  if (!fast && P.keys_in_smem) {
    int* s_cnt = reinterpret_cast<int*>(sh->top);   // top[] is not live yet: C <= 2048 ints fit
    const bool cnt_smem = C <= 2048;
    if (cnt_smem) {
      for (int c = tid; c < C; c += RPP_MERGE_NT) s_cnt[c] = scnt[c];
      __syncthreads();
    }
    for (int i = tid; i < C * M; i += RPP_MERGE_NT) {
      const int c = i / M, slot = i - c * M;
      const int n = cnt_smem ? s_cnt[c] : scnt[c];
      skeys[i] = slot < n ? make_key(key_score(sk[i]), (u32)i) : (P.combined ? 0ull : make_key(0.0f, (u32)i));
    }
    __syncthreads();
  }
  u64 KB = ~0ull;
  while (!fast && got < M) {
    const int m = P.keys_in_smem
        ? select_chunk<RPP_MERGE_NT>([&](int i) { return skeys[i]; }, C * M, KB, M - got, sh->chunk, RPP_CHUNK_CAP,
                                     &sh->sel)
        : select_chunk<RPP_MERGE_NT>(keyfn, C * M, KB, M - got, sh->chunk, RPP_CHUNK_CAP, &sh->sel);
    if (m == 0) break;
    const int take = m < M - got ? m : M - got;
    for (int i = tid; i < take; i += RPP_MERGE_NT) sh->top[got + i] = sh->chunk[i];
    got += take;
    __syncthreads();
  }
  // valid count
  if (tid == 0) sh->npos = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < got; i += RPP_MERGE_NT)
    if (P.combined || key_score(sh->top[i]) > 0.0f) ++local;  // :481-482 count(score > 0)
  if (local) atomicAdd(&sh->npos, local);
  __syncthreads();
  const int valid = sh->npos;
  if (tid == 0) P.out_valid[b] = valid;

  float4* ob = P.out_boxes + (size_t)b * M;
  float* os = P.out_scores + (size_t)b * M;
  for (int i = tid; i < M; i += RPP_MERGE_NT) {
    sh->need[i] = -1;
    if (P.combined) {
      if (i < valid) {
        const u32 flat = key_tie(sh->top[i]);
        ob[i] = clip01(P.sel_box[(size_t)b * C * M + flat]);  // clip_boxes=True (:234)
        os[i] = key_score(sh->top[i]);
        ((float*)P.out_classes)[(size_t)b * M + i] = (float)(flat / M);
      } else {
        ob[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        os[i] = 0.0f;
        ((float*)P.out_classes)[(size_t)b * M + i] = 0.0f;
      }
    } else {
      const u32 flat = key_tie(sh->top[i]);
      const int c = flat / M, slot = flat - c * M;
      os[i] = i < valid ? key_score(sh->top[i]) : -1.0f;                       // :484-486
      ((int*)P.out_classes)[(size_t)b * M + i] = i < valid ? c : -1;           // :488-490
      if (slot < scnt[c]) ob[i] = P.sel_box[(size_t)b * C * M + flat];
      else if (P.row0_mode == 1 && scnt[c] > 0) ob[i] = P.sel_box[((size_t)b * C + c) * M];  // row 0 = best kept
      else sh->need[i] = c;   // NMSV5 pads indices with 0 (:453): row 0 of this class's input list
    }
  }
  if (P.combined) return;
  __syncthreads();
  // pad boxes that are "row 0" of a class: index 0 of a dense / unfiltered input (row0_mode 0), or the best element
  // of the column when the per-class top-k ran first and the class kept nothing (row0_mode 1).
  int last_c = -1;
  float4 last_box = make_float4(0.f, 0.f, 0.f, 0.f);
  // pads have score 0 and sort after every positive score: they can only sit at positions >= valid (for a
  // negative score threshold a 0-score pad may precede a negative kept score, hence min(valid, first pad))
  const int first = P.score_nonneg ? valid : 0;
  for (int i = first; i < M; ++i) {
    const int c = sh->need[i];  // uniform across the block
    if (c < 0) continue;
    if (c != last_c) {
      u32 row = 0;
      if (P.row0_mode == 1) {
        // argmax of the column under (score desc, index asc)
        u64 best = 0ull;
        for (long r = tid; r < P.N; r += RPP_MERGE_NT) {
          const float raw = lv_val(P.lv, b, r, P.C, c);
          const u64 k = ((u64)ord_f32(raw) << 32) | (u64)(0xffffffffu - (u32)r);
          best = k > best ? k : best;
        }
        u32 cnt = 0; u64 mn = ~0ull;
        block_cnt_max_min<RPP_MERGE_NT>(cnt, best, mn, &sh->sel.bs);
        row = 0xffffffffu - (u32)best;
        if (P.is_logit) {
          // different logits can round to the same score: the reference's order is by SCORE then index
          const float raw_max = unord_f32((u32)(best >> 32));
          const float s_max = sigmoid_f32(raw_max);
          // lowest logit that still rounds to s_max (sigmoid is monotone): bisection on the ordered encoding
          u32 lo_o = ord_f32(-INFINITY), hi_o = (u32)(best >> 32);   // S(lo) < s_max (or lo = -inf), S(hi) == s_max
          if (sigmoid_f32(-INFINITY) == s_max) hi_o = lo_o;
          while (hi_o - lo_o > 1u) {
            const u32 mid = lo_o + ((hi_o - lo_o) >> 1);
            if (sigmoid_f32(unord_f32(mid)) == s_max) hi_o = mid; else lo_o = mid;
          }
          const float raw_lo = unord_f32(hi_o);
          u64 best2 = 0ull;
          for (long r = tid; r < P.N; r += RPP_MERGE_NT) {
            const float raw = lv_val(P.lv, b, r, P.C, c);
            if (raw >= raw_lo) { const u64 k = (u64)(0xffffffffu - (u32)r); best2 = k > best2 ? k : best2; }
          }
          cnt = 0; mn = ~0ull;
          block_cnt_max_min<RPP_MERGE_NT>(cnt, best2, mn, &sh->sel.bs);
          row = 0xffffffffu - (u32)best2;
        }
      }
      float4 bx;
      if (P.boxes) {
        const int qi = P.q > 1 ? (c < P.q - 1 ? c : P.q - 1) : 0;
        bx = P.boxes[((size_t)b * P.N + row) * P.q + qi];
      } else {
        bx = decode_box(lv_delta(P.lv, b, row), P.anchors[row], P.dp);
      }
      last_box = clip01(bx);
      last_c = c;
    }
    if (tid == 0) ob[i] = last_box;
  }
}

// Per-image merge of _tpu_per_class_hard_nms (postprocessing_ops.py:337-379): the C*M per-class slots — the kept
// boxes, then for a class that kept fewer than M the padded slots, which gather that class's index 0 (box AND
// score, :332-335) — go through tf.nn.top_k(M) (score desc, flat index asc) and every position whose score is not
// above the score threshold becomes -1 in all fields.  Only slots scoring above the threshold can surface, so the
// keys of the others are left out.
struct MergePaddedParams {
  int C, M;
  float score_threshold;
  const u64* sel_key; const float4* sel_box; const int* sel_cnt;
  const float* pad_score; const float4* pad_box;
  float4* out_boxes; float* out_scores; int* out_classes; int* out_valid;
};

__global__ void __launch_bounds__(RPP_MERGE_NT) merge_padded_kernel(MergePaddedParams P) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MergeShared* sh = reinterpret_cast<MergeShared*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int C = P.C, M = P.M;
  const u64* sk = P.sel_key + (size_t)b * C * M;
  const int* scnt = P.sel_cnt + (size_t)b * C;
  const float* ps = P.pad_score + (size_t)b * C;
  auto keyfn = [&](int i) -> u64 {
    const int c = i / M, slot = i - c * M;
    const float s = slot < scnt[c] ? key_score(sk[i]) : ps[c];
    return s > P.score_threshold ? make_key(s, (u32)i) : 0ull;
  };
  int got = 0;
  u64 KB = ~0ull;
  while (got < M) {
    const int m = select_chunk<RPP_MERGE_NT>(keyfn, C * M, KB, M - got, sh->chunk, RPP_CHUNK_CAP, &sh->sel);
    if (m == 0) break;
    const int take = m < M - got ? m : M - got;
    for (int i = tid; i < take; i += RPP_MERGE_NT) sh->top[got + i] = sh->chunk[i];
    got += take;
    __syncthreads();
  }
  if (tid == 0) P.out_valid[b] = got;   // :361-363: count of positions above the threshold
  for (int i = tid; i < M; i += RPP_MERGE_NT) {
    const size_t o = (size_t)b * M + i;
    if (i < got) {
      const u32 flat = key_tie(sh->top[i]);
      const int c = flat / M, slot = flat - c * M;
      P.out_boxes[o] = slot < scnt[c] ? P.sel_box[(size_t)b * C * M + flat] : P.pad_box[(size_t)b * C + c];
      P.out_scores[o] = key_score(sh->top[i]);
      P.out_classes[o] = c;
    } else {
      P.out_boxes[o] = make_float4(-1.f, -1.f, -1.f, -1.f);
      P.out_scores[o] = -1.0f;
      P.out_classes[o] = -1;
    }
  }
}

// ===============================================================================================================
// K5  Global* modes (GenerateDetections._global_nms, postprocessing_ops.py:244-286): NonMaxSuppressionV5 runs on
// the per-row maximum over classes.  rowmax_kernel reduces [B,n,C] -> [B,n] (max raw value per row; the score is
// monotone in the raw value so max score = score(max raw)); the problem kernel then runs with C = 1;
// global_out_kernel gathers boxes / classes and applies the reference's padding (score -1, class -1, box =
// boxes[0]).  The class (tf.argmax: first maximum, by SCORE) is only needed for the <= M selected
// rows, so it is resolved there.
// ===============================================================================================================
// thread-per-row variant for narrow rows (C <= 16): adjacent threads read adjacent rows
__global__ void rowmax_small_kernel(const float* __restrict__ x, size_t rows, int C, float* __restrict__ out) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(x + r * C + c));
    out[r] = m;
  }
}

__global__ void rowmax_kernel(const float* __restrict__ x, size_t rows, int C, float* __restrict__ out) {
  // one warp per row: coalesced reads of the row's C values
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < rows; r += nwarps) {
    float m = -INFINITY;
    bool any_nan = false;
    for (int c = lane; c < C; c += 32) {
      const float v = __ldg(x + r * C + c);
      any_nan |= v != v;
      m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(RPP_FULL_MASK, m, o));
    if (lane == 0) out[r] = m;
  }
}

// [B,N,C] in per-level pieces and / or 16-bit elements -> [B,N] row maxima (one warp per row)
__global__ void rowmax_levels_kernel(Levels lv, long N, int C, size_t rows, float* __restrict__ out) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < rows; r += nwarps) {
    const int b = (int)(r / (size_t)N);
    const long row = (long)(r - (size_t)b * N);
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, lv_val(lv, b, row, C, c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(RPP_FULL_MASK, m, o));
    if (lane == 0) out[r] = m;
  }
}

struct GlobalOutParams {
  int M;
  const u64* sel_key;     // [B][M]  (score | ~row)
  const float4* sel_box;  // [B][M]
  const int* sel_cnt;     // [B]
  Levels lv;              // class lookup source: logits (is_logit) or scores, [B, rows, C] (fused or per-level pieces)
  int is_logit; long n; int C;
  // rows of the NMS input that are not rows of `lv`: row j of the filtered list is anchor tie(row_keys[b][j]) / C
  // (the global filter without the row gather, rpp_global.cuh); nullptr: the NMS rows are the rows of `lv`
  const u64* row_keys; long k_rows;
  const float4* deltas; const float4* anchors; const float4* boxes; DecodeParams dp;
  float4* out_boxes; float* out_scores; long long* out_classes; int* out_valid;
  int tpu;                // _tpu_global_hard_nms (:402-431): int32 classes, -1 in every field beyond valid
};

__global__ void global_out_kernel(GlobalOutParams P) {
  pdl_enter();
  const int b = blockIdx.x;
  const int valid = P.sel_cnt[b];
  if (threadIdx.x == 0) P.out_valid[b] = valid;
  for (int i = threadIdx.x; i < P.M; i += blockDim.x) {
    const size_t o = (size_t)b * P.M + i;
    if (i < valid) {
      const u64 k = P.sel_key[o];
      u32 row = key_tie(k);
      if (P.row_keys) row = key_tie(P.row_keys[(size_t)b * P.k_rows + row]) / (u32)P.C;
      // tf.argmax over scores: first class whose SCORE equals the row maximum
      float best = -INFINITY;
      for (int c = 0; c < P.C; ++c) best = fmaxf(best, lv_val(P.lv, b, row, P.C, c));
      const float s_best = P.is_logit ? sigmoid_f32(best) : best;
      int cls = 0;
      for (int c = 0; c < P.C; ++c) {
        const float raw = lv_val(P.lv, b, row, P.C, c);
        const float s = P.is_logit ? (raw == best ? s_best : sigmoid_f32(raw)) : raw;
        if (s == s_best) { cls = c; break; }
      }
      P.out_boxes[o] = P.sel_box[o];
      P.out_scores[o] = key_score(k);
      if (P.tpu) reinterpret_cast<int*>(P.out_classes)[o] = cls; else P.out_classes[o] = cls;
    } else if (P.tpu) {
      P.out_boxes[o] = make_float4(-1.f, -1.f, -1.f, -1.f);
      P.out_scores[o] = -1.0f;
      reinterpret_cast<int*>(P.out_classes)[o] = -1;
    } else {
      // padded selected index 0 -> boxes[0] (clipped), score -1, class -1 (:258-268)
      u32 row0 = 0u;
      if (P.row_keys) {
        const u64 k0 = P.row_keys[(size_t)b * P.k_rows];
        row0 = k0 != 0ull ? key_tie(k0) / (u32)P.C : 0u;
      }
      float4 bx = P.boxes ? P.boxes[(size_t)b * P.n] : decode_box(lv_delta(P.lv, b, row0), P.anchors[row0], P.dp);
      P.out_boxes[o] = clip01(bx);
      P.out_scores[o] = -1.0f;
      P.out_classes[o] = -1;
    }
  }
}

// ===============================================================================================================
// K8  EfficientNMS_TRT-compatible entry (the node the reference appends for export mode onnx_tensorrt,
// onnx_utils.py:13-85, with the attributes it sets: score_activation = sigmoid, box_coding = 1 (centre-size, decoded
// against the anchor input), background_class = -1, class-aware suppression).  The emission kernels deliver, per
// image, the RPP_EFFNMS_SELECTED best (anchor, class) pairs sorted by (score desc, flat index asc); this kernel walks
// them greedily — a candidate is dropped when a kept box of the SAME class overlaps it by more than iou_threshold —
// until max_output_boxes are kept, and writes the plugin's four outputs (zero-filled beyond the count).
// Same tile bit-mask / bit-chain structure as hard_nms_consume.
// ===============================================================================================================
#define RPP_EFFNMS_SELECTED 4096

struct EffNmsParams {
  const u64* emit_key; long k;        // [B][k] sorted keys (score bits | ~flat index), flat = anchor * C + class
  const float4* deltas;               // [B][N] raw boxes (dx, dy, dw, dh)
  const float4* anchors;              // [N] (cx, cy, w, h)
  long N; int C;
  float score_threshold, iou_threshold;
  int M;
  int* out_valid; float4* out_boxes; float* out_scores; int* out_classes;
};

__global__ void __launch_bounds__(RPP_NMS_NT) effnms_kernel(EffNmsParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* kbox = reinterpret_cast<float4*>(smem_raw);              // [M] kept boxes (corner coding)
  float* karea = reinterpret_cast<float*>(kbox + P.M);             // [M]
  int* kcls = reinterpret_cast<int*>(karea + P.M);                 // [M]
  __shared__ float4 cbox[RPP_NMS_NT];
  __shared__ float carea[RPP_NMS_NT];
  __shared__ int ccls[RPP_NMS_NT];
  __shared__ int s_nkept, s_slot[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float thr = P.iou_threshold;
  for (int i = tid; i < P.M; i += RPP_NMS_NT) {   // the plugin clears its outputs first
    const size_t o = (size_t)b * P.M + i;
    P.out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    P.out_scores[o] = 0.0f;
    P.out_classes[o] = 0;
  }
  if (tid == 0) s_nkept = 0;
  __syncthreads();
  bool full = false;
  for (long g0 = 0; g0 < P.k; g0 += RPP_NMS_NT) {
    const u64 key = g0 + tid < P.k ? P.emit_key[(size_t)b * P.k + g0 + tid] : 0ull;
    const float score = key_score(key);
    bool alive = key != 0ull && score >= P.score_threshold;        // sorted: the candidates are a prefix
    const int gcount = __syncthreads_count(alive);
    if (gcount == 0) break;
    float4 bx = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY), corner = bx;
    float area = 0.0f;
    int cls = -1;
    if (alive) {
      const u32 flat = key_tie(key);
      const u32 row = flat / (u32)P.C;
      cls = (int)(flat - row * (u32)P.C);
      // centre-size decode against the anchor, no variance scaling, no normalisation
      const float4 d = P.deltas[(size_t)b * P.N + row];
      const float4 a = P.anchors[row];
      const float cx = __fadd_rn(__fmul_rn(d.x, a.z), a.x);
      const float cy = __fadd_rn(__fmul_rn(d.y, a.w), a.y);
      const float hw = __fmul_rn(__fmul_rn(a.z, exp_f32(d.z)), 0.5f);
      const float hh = __fmul_rn(__fmul_rn(a.w, exp_f32(d.w)), 0.5f);
      corner = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
      const float w = __fsub_rn(corner.z, corner.x), hgt = __fsub_rn(corner.w, corner.y);
      if (w > 0.0f && hgt > 0.0f) { area = __fmul_rn(w, hgt); bx = corner; }
    }
    cbox[tid] = bx; carea[tid] = area; ccls[tid] = cls;
    int nk = s_nkept;   // read before the barrier; inside the tile loop the count travels through s_slot[]
    __syncthreads();
    u32 rowm = 0u;
    {
      const int tbase = warp * 32;
      const int tcount = gcount - tbase < 32 ? gcount - tbase : 32;
      for (int j = 0; j < tcount - 1; ++j)
        if (j < lane && lane < tcount && ccls[tbase + j] == cls && iou_gt(bx, area, cbox[tbase + j], carea[tbase + j], thr))
          rowm |= 1u << j;
    }
    int tested = 0;
    const int ntiles = (gcount + 31) >> 5;
    for (int tile = 0; tile < ntiles; ++tile) {
      if (alive && warp >= tile) {
        for (int q = tested; q < nk; ++q)
          if (kcls[q] == cls && iou_gt(bx, area, kbox[q], karea[q], thr)) { alive = false; break; }
      }
      tested = nk;
      if (warp == tile) {
        const u32 cand_bits = __ballot_sync(RPP_FULL_MASK, alive);
        u32 kept_bits = 0u;
#pragma unroll
        for (int l = 0; l < 32; ++l) {
          const u32 r = __shfl_sync(RPP_FULL_MASK, rowm, l);
          if (((cand_bits >> l) & 1u) && (r & kept_bits) == 0u) kept_bits |= 1u << l;
        }
        int nnew = __popc(kept_bits);
        const int room = P.M - nk;
        while (nnew > room) {
          kept_bits &= ~(1u << (31 - __clz(kept_bits)));
          --nnew;
        }
        if ((kept_bits >> lane) & 1u) {
          const int pos = nk + __popc(kept_bits & ((1u << lane) - 1u));
          kbox[pos] = bx; karea[pos] = area; kcls[pos] = cls;
          const size_t o = (size_t)b * P.M + pos;
          // outputs use the input's box coding (centre-size), rebuilt from the corner box
          const float w = __fsub_rn(corner.z, corner.x), hgt = __fsub_rn(corner.w, corner.y);
          P.out_boxes[o] = make_float4(__fadd_rn(corner.x, __fmul_rn(0.5f, w)), __fadd_rn(corner.y, __fmul_rn(0.5f, hgt)),
                                       w, hgt);
          P.out_scores[o] = score;
          P.out_classes[o] = cls;
        }
        if (lane == 0) {
          s_slot[(tile + 1) & 1] = nk + nnew;
          s_nkept = nk + nnew;
        }
      }
      __syncthreads();
      nk = s_slot[(tile + 1) & 1];
      if (nk >= P.M) { full = true; break; }
    }
    if (full || gcount < RPP_NMS_NT) break;
  }
  __syncthreads();
  if (tid == 0) P.out_valid[b] = s_nkept;
}

// ===============================================================================================================
// K6  FilterTopKDetections outputs (postprocessing_ops.py:128-161) from the emitted sorted keys.
// ===============================================================================================================
// per class: scores_out [B,k,C], boxes_out [B,k,C,4], idx_out [B,C,k]
// (K_out / j_off / idx_off: the outputs are rows [j_off, j_off + k) of tensors with K_out rows per image, and the
// indices are offset by idx_off — the per-level segments of rpp_topk_levels; K_out = k, 0, 0 otherwise)
__global__ void topk_gather_per_class_kernel(const u64* __restrict__ emit_key /*[B*C][k]*/, const float4* __restrict__ boxes,
                                             int B, long n, int C, long k, float* __restrict__ scores_out,
                                             float4* __restrict__ boxes_out, int* __restrict__ idx_out, long K_out,
                                             long j_off, int idx_off) {
  const size_t tot = (size_t)B * k * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t bj = e / C;
    const long j = (long)(bj % k);
    const int b = (int)(bj / k);
    const u64 key = emit_key[((size_t)b * C + c) * k + j];
    const u32 row = key_tie(key);
    const size_t eo = ((size_t)b * K_out + j_off + j) * C + c;
    scores_out[eo] = key_score(key);
    boxes_out[eo] = boxes[(size_t)b * n + row];
    if (idx_out) idx_out[((size_t)b * C + c) * K_out + j_off + j] = (int)row + idx_off;
  }
}

// global: emitted keys over the flat [n*C] axis; scores_out [B,k,C] = whole rows, boxes_out [B,k,4]
__global__ void topk_gather_global_kernel(const u64* __restrict__ emit_key /*[B][k]*/, const float* __restrict__ scores,
                                          const float4* __restrict__ boxes, int B, long n, int C, long k,
                                          float* __restrict__ scores_out, float4* __restrict__ boxes_out,
                                          int* __restrict__ idx_out, long K_out, long j_off, int idx_off) {
  const size_t tot = (size_t)B * k * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t bj = e / C;
    const int b = (int)(bj / k);
    const u32 flat = key_tie(emit_key[bj]);
    const u32 a = flat / (u32)C;  // indices // num_classes (:156)
    const size_t ro = (size_t)b * K_out + j_off + (bj - (size_t)b * k);
    scores_out[ro * C + c] = scores[((size_t)b * n + a) * C + c];
    if (c == 0) {
      boxes_out[ro] = boxes[(size_t)b * n + a];
      if (idx_out) idx_out[ro] = (int)flat + idx_off * C;
    }
  }
}

// fused global filter: rows selected on raw logits -> materialise the reference's intermediates
// scores [B,k,C] = sigmoid(logit rows), boxes [B,k,4] = decoded anchors (TransformBoxesAndScores on k rows only)
__global__ void fused_global_rows_kernel(const u64* __restrict__ emit_key /*[B][k]*/, Levels lv /*logits + deltas*/,
                                         const float4* __restrict__ anchors, DecodeParams dp, int B, long N, int C,
                                         long k, int apply_sigmoid, float* __restrict__ scores_out,
                                         float4* __restrict__ boxes_out) {
  if (C < 16) {   // narrow rows: one thread per element
    const size_t tot = (size_t)B * k * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
      const int c = (int)(e % C);
      const size_t bj = e / C;
      const int b = (int)(bj / k);
      const u32 a = key_tie(emit_key[bj]) / (u32)C;
      const float raw = lv_val(lv, b, a, C, c);
      scores_out[e] = apply_sigmoid ? sigmoid_f32(raw) : raw;
      if (c == 0) boxes_out[bj] = decode_box(lv_delta(lv, b, a), anchors[a], dp);
    }
    return;
  }
  // a group of tpr = min(32, pow2 >= C) threads per selected row: the row index arithmetic (64-bit divisions) is done
  // once per row, the C values of the row are read and written coalesced
  const size_t rows = (size_t)B * k;
  int tpr = 1;
  while (tpr < C && tpr < 32) tpr <<= 1;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = (int)(gtid & (size_t)(tpr - 1));
  const size_t warp = gtid / tpr;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) / tpr;
  for (size_t bj = warp; bj < rows; bj += nwarps) {
    const int b = (int)(bj / k);
    const u32 a = key_tie(emit_key[bj]) / (u32)C;
    float* dst = scores_out + bj * C;
    for (int c = lane; c < C; c += tpr) {
      const float raw = lv_val(lv, b, a, C, c);
      dst[c] = apply_sigmoid ? sigmoid_f32(raw) : raw;
    }
    if (lane == 0) boxes_out[bj] = decode_box(lv_delta(lv, b, a), anchors[a], dp);
  }
}

// ===============================================================================================================
// K7  COCO post-formatting epilogue — COCOEvaluator.accumulate_results (eval/coco_evaluator.py:111-134): slice by
// valid_detections, boxes /= (resize_scale / input_shape) tiled to 4 (the reference divides [x1,y1,x2,y2] by the
// [H,W,H,W]-ordered scale), np.int32 truncation, x2y2 -> wh, optional class-id remap; rows are compacted in image
// order so one small device->host copy replaces the per-image numpy loop.
// ===============================================================================================================
struct CocoParams {
  const float4* boxes; const float* scores; const void* classes; const int* valid;
  int class_kind;            // 0 f32, 1 i64, 2 i32 (per NMS mode)
  int B, M;
  const float* resize_scale; // [B,2] or nullptr (rescale_detections=False)
  float in_h, in_w;          // input.input_shape
  const int* class_map;      // [num_classes] or nullptr
  int num_classes;
  int4* bbox_out; int* category_out; float* score_out; int* image_out; int* total_out;
};

__global__ void coco_format_kernel(CocoParams P) {
  const int b = blockIdx.x;
  __shared__ int s_off;
  if (threadIdx.x == 0) {
    int off = 0;
    for (int i = 0; i < b; ++i) off += max(0, min(P.valid[i], P.M));
    s_off = off;
    if (b == P.B - 1) *P.total_out = off + max(0, min(P.valid[b], P.M));
  }
  __syncthreads();
  const int v = max(0, min(P.valid[b], P.M));
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
  if (P.resize_scale) {
    const float s0 = __fdiv_rn(P.resize_scale[2 * b + 0], P.in_h), s1 = __fdiv_rn(P.resize_scale[2 * b + 1], P.in_w);
    sc = make_float4(s0, s1, s0, s1);
  }
  for (int i = threadIdx.x; i < v; i += blockDim.x) {
    const size_t o = (size_t)b * P.M + i;
    float4 bx = P.boxes[o];
    if (P.resize_scale)
      bx = make_float4(__fdiv_rn(bx.x, sc.x), __fdiv_rn(bx.y, sc.y), __fdiv_rn(bx.z, sc.z), __fdiv_rn(bx.w, sc.w));
    const int x1 = (int)bx.x, y1 = (int)bx.y, x2 = (int)bx.z, y2 = (int)bx.w;   // np.int32: truncation toward zero
    int cls;
    if (P.class_kind == 0) cls = (int)((const float*)P.classes)[o];
    else if (P.class_kind == 1) cls = (int)((const long long*)P.classes)[o];
    else cls = ((const int*)P.classes)[o];
    if (P.class_map && cls >= 0 && cls < P.num_classes) cls = P.class_map[cls];
    const size_t r = (size_t)s_off + i;
    P.bbox_out[r] = make_int4(x1, y1, x2 - x1, y2 - y1);
    P.category_out[r] = cls;
    P.score_out[r] = P.scores[o];
    P.image_out[r] = b;
  }
}
