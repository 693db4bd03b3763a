// rpp_global.cuh — K9: the Global* modes behind the global pre-NMS filter (GlobalSoftNMS / GlobalHardNMS with
// inference.filter_per_class = false: FilterTopKDetections._filter_global, postprocessing_ops.py:149-161, followed by
// GenerateDetections._global_nms, :244-286) without ever gathering the k selected rows.
// Part of the retinapost kernel set; included by rpp_kernels.cuh (one translation unit: rpp_api.cu).
#pragma once
#include "rpp_kernels.cuh"

// ===============================================================================================================
// The reference gathers, for each of the k best (anchor, class) pairs, the WHOLE score row of that anchor and its box
// (:156-159; an anchor with several classes in the top k appears several times), and _global_nms then
// runs NonMaxSuppressionV5 on the row maxima.  Nothing of that needs the rows themselves:
//   * the maximum of row a is the score of the FIRST pair of anchor a in the sorted top-k list: if (a, c) is in the
//     top k, so is (a, argmax_c) — it scores at least as much and, on equal scores, the lower flat index wins;
//   * tf.argmax of the row = the class of that first pair, for the same reason (resolved by global_out_kernel /
//     the epilogue of global_soft_kernel from the <= M selected rows' logits, on scores).
// global_rows_kernel therefore turns the sorted keys emit_key[b][j] (score bits | ~flat index) into the NMS input
//   mraw[b][j]      = max-class score of filtered row j   (= score of the first occurrence of its anchor)
// (the row's box is decoded from its anchor by whoever consumes the row) and, for the soft-NMS kernel, the candidates in NonMaxSuppressionV5's order (score desc, row index asc):
//   skey[b][0..ns)  rows that are the first occurrence of their anchor — already in order, compacted;
//   dkey[b][0..nd)  duplicate rows (score = their anchor's maximum, which is larger than their own pair's): unsorted.
// `first` is scratch [B][N] preset to 0xffffffff: first[b][a] = smallest j whose anchor is a (atomicMin).
// ===============================================================================================================
#define RPP_GROWS_NT 512

struct GlobalRowsParams {
  const u64* emit_key;   // [B][k]
  long k;
  int C;
  long N;
  u32* first;            // [B][N]
  float* mraw;           // [B][k]
  u64* skey;             // [B][k] or nullptr
  u64* dkey;             // [B][k]
  int* sd_cnt;           // [B][2] = {ns, nd}
  float score_threshold;
  const int* skip;       // [B] or nullptr: 2 = global_top_direct_kernel already did the image
  int init_first;        // the block presets its image's `first` itself (no memset ahead of the kernel)
  int n_loop;            // > 0: persistent blocks walk images [0, n_loop)
};

// candidate key of the Global* NMS: (score, row index j) with the row's payload slot free in the low bits
__device__ __forceinline__ u64 grow_key(float score, u32 j) { return make_key(score, j); }

#define RPP_GROWS_RPT 8   // consecutive rows per thread

__device__ __forceinline__ void global_rows_body(const GlobalRowsParams& P, const int b) {
  __shared__ int s_wsum[RPP_GROWS_NT / 32];
  __shared__ int s_run, s_nd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64* ek = P.emit_key + (size_t)b * P.k;
  u32* fr = P.first + (size_t)b * P.N;
  const long span = (long)RPP_GROWS_NT * RPP_GROWS_RPT;
  if (P.skip && P.skip[b] == 2) return;
  if (P.init_first) {
    for (long i = tid; i < P.N; i += RPP_GROWS_NT) fr[i] = 0xffffffffu;
    __syncthreads();
  }
  // pass A: first[anchor] = lowest row index of the anchor (loads first, then the reductions: nothing waits)
  for (long j0 = 0; j0 < P.k; j0 += span) {
    u64 key[RPP_GROWS_RPT];
#pragma unroll
    for (int i = 0; i < RPP_GROWS_RPT; ++i) {
      const long j = j0 + (long)tid * RPP_GROWS_RPT + i;
      key[i] = j < P.k ? ek[j] : 0ull;
    }
#pragma unroll
    for (int i = 0; i < RPP_GROWS_RPT; ++i)
      if (key[i] != 0ull) atomicMin(&fr[key_tie(key[i]) / (u32)P.C], (u32)(j0 + (long)tid * RPP_GROWS_RPT + i));
  }
  if (tid == 0) { s_run = 0; s_nd = 0; }
  __syncthreads();   // the block's own global atomics are visible to the block after the barrier (read with ld.cg)
  // pass B: thread t owns RPP_GROWS_RPT consecutive rows, so the first-occurrence rows compact in order with one
  // block-wide prefix sum per span
  for (long j0 = 0; j0 < P.k; j0 += span) {
    const long jb = j0 + (long)tid * RPP_GROWS_RPT;
    u64 key[RPP_GROWS_RPT];
    u32 f[RPP_GROWS_RPT];
    float s[RPP_GROWS_RPT];
#pragma unroll
    for (int i = 0; i < RPP_GROWS_RPT; ++i) key[i] = jb + i < P.k ? ek[jb + i] : 0ull;
#pragma unroll
    for (int i = 0; i < RPP_GROWS_RPT; ++i)
      f[i] = key[i] != 0ull ? __ldcg(&fr[key_tie(key[i]) / (u32)P.C]) : 0xffffffffu;
    int n_first = 0;
    u32 firsts = 0u, dups = 0u;
#pragma unroll
    for (int i = 0; i < RPP_GROWS_RPT; ++i) {
      s[i] = -INFINITY;
      if (key[i] != 0ull) {
        const bool is_first = f[i] == (u32)(jb + i);
        s[i] = key_score(is_first ? key[i] : ek[f[i]]);
        // NonMaxSuppressionV5 only ever sees candidates above the score threshold (A.2)
        if (s[i] > P.score_threshold) {
          if (is_first) { firsts |= 1u << i; ++n_first; } else { dups |= 1u << i; }
        }
      }
      if (jb + i < P.k) P.mraw[(size_t)b * P.k + jb + i] = s[i];
    }
    if (P.skey) {
      int incl = n_first;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(RPP_FULL_MASK, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int before = s_run + incl - n_first;
      for (int w = 0; w < warp; ++w) before += s_wsum[w];
#pragma unroll
      for (int i = 0; i < RPP_GROWS_RPT; ++i) {
        if ((firsts >> i) & 1u) P.skey[(size_t)b * P.k + before++] = grow_key(s[i], (u32)(jb + i));
        if ((dups >> i) & 1u) P.dkey[(size_t)b * P.k + atomicAdd(&s_nd, 1)] = grow_key(s[i], (u32)(jb + i));
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < RPP_GROWS_NT / 32; ++w) tot += s_wsum[w];
        s_run += tot;
      }
      __syncthreads();
    }
  }
  if (P.skey && tid == 0) { P.sd_cnt[2 * b] = s_run; P.sd_cnt[2 * b + 1] = s_nd; }
}

__global__ void __launch_bounds__(RPP_GROWS_NT) global_rows_kernel(GlobalRowsParams P) {
  pdl_enter();
  if (P.n_loop > 0) {   // a few persistent blocks over images that are almost all done (see GlobalRowsParams.skip)
    for (int b = blockIdx.x; b < P.n_loop; b += gridDim.x) {
      global_rows_body(P, b);
      __syncthreads();
    }
  } else {
    global_rows_body(P, blockIdx.x);
  }
}

// ===============================================================================================================
// global_top_kernel — GlobalHardNMS behind the global filter, non-TPU branch.  The reference passes iou_threshold = 1.0
// to NonMaxSuppressionV5 there (postprocessing_ops.py:253), and an IoU computed as
// inter / (area_a + area_b - inter) in IEEE arithmetic never exceeds 1 (inter <= min(area_a, area_b), and rounding is
// monotone), so nothing is ever suppressed: the output is the first max_detections rows above the score threshold in
// (row maximum desc, row index asc) order — duplicates of an anchor included (B12).  Candidates: the first M
// first-occurrence rows (already in order) and every duplicate row; one selection of the M best.
// ===============================================================================================================
#define RPP_GTOP_NT 256

struct GlobalTopParams {
  long k; int M;
  const u64* skey; const u64* dkey; const int* sd_cnt;
  const u64* emit_key; Levels lv; int C; long N;
  const float4* anchors; DecodeParams dp;
  float4* out_boxes; float* out_scores; long long* out_classes; int* out_valid;
  const int* skip;       // [B] or nullptr: 2 = global_top_direct_kernel already did the image
  int n_loop;            // > 0: persistent blocks walk images [0, n_loop)
};

struct GlobalTopShared {
  SelectScratch<RPP_GTOP_NT> sel;
  u64 chunk[1024];
  u64 top[1024];
};

__device__ __forceinline__ void global_top_body(const GlobalTopParams& P, const int b) {
  __shared__ GlobalTopShared sh;
  const int tid = threadIdx.x, lane = tid & 31;
  if (P.skip && P.skip[b] == 2) return;
  const int ns = P.sd_cnt[2 * b], nd = P.sd_cnt[2 * b + 1];
  const int n1 = ns < P.M ? ns : P.M;
  const u64* sk = P.skey + (size_t)b * P.k;
  const u64* dk = P.dkey + (size_t)b * P.k;
  u64 KB = ~0ull;
  int got = 0;
  while (got < P.M) {   // (a chunk may come back shorter than asked for: the radix cut is exact, not its size)
    const int m = select_chunk<RPP_GTOP_NT>([&](int i) { return i < n1 ? sk[i] : dk[i - n1]; }, n1 + nd, KB,
                                            P.M - got, sh.chunk, 1024, &sh.sel);
    if (m == 0) break;
    const int take = m < P.M - got ? m : P.M - got;
    for (int i = tid; i < take; i += RPP_GTOP_NT) sh.top[got + i] = sh.chunk[i];
    got += take;
    __syncthreads();
  }
  const int valid = got;
  if (tid == 0) P.out_valid[b] = valid;
  for (int i = tid >> 5; i < P.M; i += RPP_GTOP_NT / 32) {   // one warp per output row (as in global_soft_kernel)
    const size_t o = (size_t)b * P.M + i;
    if (i < valid) {
      const u64 key = sh.top[i];
      const u32 j = key_tie(key);
      const u32 a = key_tie(P.emit_key[(size_t)b * P.k + j]) / (u32)P.C;
      float best = -INFINITY;
      int cls = 0x7fffffff;
      for (int c = lane; c < P.C; c += 32) {
        const float raw = lv_val(P.lv, b, a, P.C, c);
        if (raw > best) { best = raw; cls = c; }
      }
#pragma unroll
      for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const float ob = __shfl_xor_sync(RPP_FULL_MASK, best, ofs);
        const int oc = __shfl_xor_sync(RPP_FULL_MASK, cls, ofs);
        if (ob > best || (ob == best && oc < cls)) { best = ob; cls = oc; }
      }
      bool near = false;
      for (int c = lane; c < cls; c += 32) {
        const float raw = lv_val(P.lv, b, a, P.C, c);
        near = near || raw > best - 1.0f || best > 15.0f || best < -80.0f;
      }
      if (__any_sync(RPP_FULL_MASK, near)) {
        const float s_best = sigmoid_f32(best);
        int c2 = 0x7fffffff;
        for (int c = lane; c < cls; c += 32)
          if (sigmoid_f32(lv_val(P.lv, b, a, P.C, c)) == s_best && c < c2) c2 = c;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) c2 = min(c2, __shfl_xor_sync(RPP_FULL_MASK, c2, ofs));
        if (c2 < cls) cls = c2;
      }
      if (lane == 0) {
        P.out_boxes[o] = clip01(decode_box(lv_delta(P.lv, b, a), P.anchors[a], P.dp));
        P.out_scores[o] = key_score(key);
        P.out_classes[o] = cls;
      }
    } else if (lane == 0) {
      // padded selected index 0 -> boxes[0] (clipped), score -1, class -1 (:258-268)
      const u64 k0 = P.emit_key[(size_t)b * P.k];
      const u32 a0 = k0 != 0ull ? key_tie(k0) / (u32)P.C : 0u;
      P.out_boxes[o] = clip01(decode_box(lv_delta(P.lv, b, a0), P.anchors[a0], P.dp));
      P.out_scores[o] = -1.0f;
      P.out_classes[o] = -1;
    }
  }
}

__global__ void __launch_bounds__(RPP_GTOP_NT) global_top_kernel(GlobalTopParams P) {
  pdl_enter();
  if (P.n_loop > 0) {
    for (int b = blockIdx.x; b < P.n_loop; b += gridDim.x) {
      global_top_body(P, b);
      __syncthreads();
    }
  } else {
    global_top_body(P, blockIdx.x);
  }
}

// ===============================================================================================================
// global_top_direct_kernel — the whole GlobalHardNMS-behind-the-global-filter stage of one image in ONE block, straight
// from the image's candidate list: no sorted top-k emission, no row resolution.  (configs[4]: the 8 192-key sort was
// half of the step.)  With nothing suppressed (see global_top_kernel) the detections are the first M rows in
// (row maximum desc, row index asc) order, and a row's index is the rank of its own (anchor, class) pair in the top k.
// Everything is selected on RAW keys (logit bits | ~flat index; the score is a monotone function of the logit) and
// the binary64 sigmoid is evaluated for a few hundred elements per image instead of the ~10^4 of the list:
//   1. the k-th best raw key (ONE shared-memory histogram over a linear map of [min key, max key] onto 1 024 bins,
//      shared with the cut of step 2, then counting inside the key's bin) and the interval
//      [x_lo, x_hi] of logits whose score equals the score s_k of that key (found with the sigmoid itself): a pair is
//      in the top k  <=>  logit > x_hi, or logit in [x_lo, x_hi] and flat index <= idx*, where idx* is the
//      (k - #{logit > x_hi})-th smallest index of the tie group E = {logit in [x_lo, x_hi]} — TopKV2's order
//      (score desc, index asc) exactly, also when distinct logits round to one score;
//   2. the M + 32 best raw keys, scored and ranked by (score, index); those strictly above the score of the raw cut
//      are complete (edge rule), and an output row's anchor has its best pair among the M best pairs (the pairs ahead
//      of an anchor's best pair all precede that pair's row in the output order);
//   3. for each distinct anchor among them, all C classes are looked up in the logits: the pairs in the top k are
//      exactly the anchor's rows (its duplicates, B12), each keyed (anchor's best score, own pair key);
//   4. the M best of those candidates (rank by counting: a few hundred at most) are written out.
// A block that cannot serve its image (list short beyond the in-block re-collection, tie group or candidate overflow,
// a tie group that reaches below the list's threshold) leaves emit_done = 0 and the emission / rows / top kernels that
// follow do the image; they skip the images marked 2.
// ===============================================================================================================
#define RPP_GTD_CAND 3072
#define RPP_GTD_TIES 1024

struct GlobalTopDirectParams {
  Levels src;              // the [B, N, C] logits + deltas (class lookups, boxes)
  int C; long N; int M;
  const float4* anchors; DecodeParams dp;
  float score_threshold;   // of the NMS (the emission problem's own threshold is -inf: tf.nn.top_k has none)
  int debug;               // env RPP_GTD_DEBUG: blocks 0 and 300 print their per-phase cycle counts
  float4* out_boxes; float* out_scores; long long* out_classes; int* out_valid;
};

// Largest d >= 0 such that every ordered float encoding in [o, o + dir * d] has the score `s` (dir = +1 / -1): the
// score is monotone in the logit, so the predicate is a prefix; galloping 32-ary search, one sigmoid per lane and round.
// One warp; o is the encoding of a finite logit whose score is s.
__device__ u32 gtd_tie_extent(u32 o, float s, int dir) {
  const int lane = threadIdx.x & 31;
  const u32 room = dir > 0 ? ord_f32(INFINITY) - o : o - ord_f32(-INFINITY);   // stay inside [-inf, +inf]
  u32 base = 0u;       // known: predicate holds at base
  u32 step = 1u;
  for (;;) {           // gallop: find a step at which the 32 probes base + step * (lane + 1) do not all hold
    const u64 d = (u64)base + (u64)step * (u32)(lane + 1);
    const bool ok = d <= (u64)room && sigmoid_f32(unord_f32(dir > 0 ? o + (u32)d : o - (u32)d)) == s;
    const u32 m = __ballot_sync(RPP_FULL_MASK, ok);
    const int n_ok = m == 0xffffffffu ? 32 : __ffs(~m) - 1;   // probes are increasing: the holds form a prefix
    base += step * (u32)n_ok;
    if (n_ok == 32) {
      if (step > (1u << 24)) return 0xffffffffu;   // a saturated score (0 or 1): millions of logits tie -> caller gives up
      step <<= 5;
      continue;
    }
    // the extent lies in [base, base + step): refine with smaller steps
    while (step > 1u) {
      step >>= 5;
      if (step == 0u) step = 1u;
      const u64 d2 = (u64)base + (u64)step * (u32)(lane + 1);
      const bool ok2 = d2 <= (u64)room && sigmoid_f32(unord_f32(dir > 0 ? o + (u32)d2 : o - (u32)d2)) == s;
      const u32 m2 = __ballot_sync(RPP_FULL_MASK, ok2);
      // 31 probes lie strictly inside the bracket; the 32nd is its (failing) upper end
      const int n2 = m2 == 0xffffffffu ? 31 : __ffs(~m2) - 1;
      base += step * (u32)(n2 > 31 ? 31 : n2);
    }
    return base;
  }
}

static_assert(RPP_RADIX_BINS == RPP_EMIT_NT, "global_top_direct_kernel: one histogram bin per thread");
static_assert(RPP_EMIT_CHUNK * 8 >= (1024 + 1024) * 8 + RPP_GTD_CAND * 12 + RPP_GTD_TIES * 4,
              "global_top_direct_kernel: scratch layout");

struct GtdState {
  int ncand, over, fail, ne, ngt, nvalid, ntop, nkbin, nrows;
  u32 o_lo, o_hi, idx_star;
  u32 d_k, rk, d_m;
  u64 kraw;
  u32 min_o;
};

__host__ __device__ static inline size_t gtd_shared_bytes() { return sizeof(EmitShared) + 1024 * sizeof(float4); }

__global__ void __launch_bounds__(RPP_EMIT_NT) global_top_direct_kernel(ColProblemParams P, GlobalTopDirectParams G) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EmitShared* sh = reinterpret_cast<EmitShared*>(smem_raw);
  float4* tbox = reinterpret_cast<float4*>(smem_raw + sizeof(EmitShared));   // [1024] boxes of the best pairs' anchors
  __shared__ GtdState st;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;   // C == 1: one emission problem per image
  const size_t p = blockIdx.x;
  u64* topraw = sh->chunk;                   // [1024] the best raw keys, then their scored keys, then per-pair scratch
  u64* top = sh->chunk + 1024;               // [1024] keys of the k-th key's bin, then: the M best pairs as
                                             //        (score, ~flat index) keys, sorted
  u64* cown = sh->chunk + 2048;              // [RPP_GTD_CAND] candidate rows: own pair key ...
  u32* cfirst = reinterpret_cast<u32*>(sh->chunk + 2048 + RPP_GTD_CAND);   // ... and the slot of the anchor's best pair in top[]
  u32* eidx = cfirst + RPP_GTD_CAND;         // [RPP_GTD_TIES] flat indices of the tie group E
  u32* hist = sh->sel.hist;
  if (tid == 0) {
    P.emit_done[p] = 0;
    st.ncand = 0; st.over = 0; st.fail = 0; st.ne = 0; st.ngt = 0; st.nvalid = 0; st.ntop = 0; st.nkbin = 0; st.nrows = 0;
    st.min_o = 0xffffffffu;
  }
  long long tdbg[8];
  int ndbg = 0;
#define GTD_T() do { if (G.debug) tdbg[ndbg++] = clock64(); } while (0)
  GTD_T();
  // ---- 0. raw keys of the list (or of an in-block re-collection of the column) -------------------------------
  u32 o_complete;   // the keys in sh->keys are ALL elements of the column whose ordered logit is >= o_complete
  u64 kmin, kmax;
  const int n_keys = emit_prepare_raw(P, sh, p, b, 0, o_complete, kmin, kmax);
  if (n_keys < 0) return;
  GTD_T();
  // ---- 1. ONE histogram over a linear map of [min key, max key] onto 1024 bins serves both selections: the bin of
  //         the k-th best key (refined exactly inside the bin) and a cut that holds the M + 32 best keys.
  //         (An MSB-first digit would put the whole list into a dozen bins: the logits of a list share their exponent.)
  const int Mp = (long)G.M < P.k_lim ? G.M : (int)P.k_lim;
  const u32 want = (u32)(Mp + 32 < 1024 ? Mp + 32 : 1024);
  u64 mn = kmin, mx = kmax;   // (per-thread extrema from the key build)
  {
    u32 cnt = 1;
    block_cnt_max_min<RPP_EMIT_NT>(cnt, mx, mn, &sh->sel.bs);
  }
  const u64 range = mx - mn;
  const u64 q = range / 1024ull + 1ull;              // bin = floor((key - mn) / q) < 1024 (any monotone map would do)
  const u64 rq = q > 1ull ? ~0ull / q : 0ull;        // floor((2^64 - 1) / q): umulhi(x, rq) <= x / q, monotone in x
  auto bin_of = [&](u64 k) -> u32 { return q > 1ull ? (u32)__umul64hi(k - mn, rq) : (u32)(k - mn); };
  hist[tid] = 0u;
  __syncthreads();
  for (int i = tid; i < n_keys; i += RPP_EMIT_NT) atomicAdd(&hist[bin_of(sh->keys[i])], 1u);
  __syncthreads();
  {
    // thread t owns bin t: inclusive suffix count over the bins >= t
    const u32 h = hist[tid];
    u32 v = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_down_sync(RPP_FULL_MASK, v, o);
      if (lane + o < 32) v += t;
    }
    if (lane == 0) sh->sel.part[warp] = v;
    __syncthreads();
    if (warp == 0) {   // exclusive suffix sums of the 32 warp totals
      const u32 tot = sh->sel.part[lane];
      u32 sfx = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_down_sync(RPP_FULL_MASK, sfx, o);
        if (lane + o < 32) sfx += t;
      }
      sh->sel.part[32 + lane] = sfx - tot;
    }
    __syncthreads();
    const u32 above = v - h + sh->sel.part[32 + warp];
    const u32 kth = (u32)P.k_lim;
    if (above < kth && above + h >= kth) { st.d_k = (u32)tid; st.rk = kth - above; if (h > 1024u) st.fail = 1; }
    if (above < want && above + h >= want) { st.d_m = (u32)tid; if (above + h > 1024u) st.fail = 1; }
    if (tid == 0 && above + h < want) st.d_m = 0u;   // fewer keys than wanted: all of them
  }
  __syncthreads();
  if (st.fail) return;
  const u32 d_k = st.d_k, d_m = st.d_m, rk = st.rk;
  for (int i = tid; i < n_keys; i += RPP_EMIT_NT) {
    const u64 k = sh->keys[i];
    const u32 bin = bin_of(k);
    if (bin >= d_m) { topraw[atomicAdd(&st.ntop, 1)] = k; atomicMin(&st.min_o, (u32)(k >> 32)); }
    if (bin == d_k) top[atomicAdd(&st.nkbin, 1)] = k;
  }
  __syncthreads();
  const int m_top = st.ntop, nkbin = st.nkbin;
  for (int i = tid; i < nkbin; i += RPP_EMIT_NT) {   // the rk-th largest key of the bin (rank by counting)
    const u64 k = top[i];
    u32 rank = 0;
    for (int j = 0; j < nkbin; ++j) rank += top[j] > k;
    if (rank == rk - 1u) st.kraw = k;
  }
  __syncthreads();
  const u32 o_k = (u32)(st.kraw >> 32);
  GTD_T();
  // ---- 2. warps 0 / 1: extent of the tie group of s_k above / below x_k; everybody else: score the best raw keys
  const float s_k = sigmoid_f32(unord_f32(o_k));
  if (warp == 0) {
    const u32 d = gtd_tie_extent(o_k, s_k, +1);
    if (lane == 0) { if (d == 0xffffffffu) st.fail = 1; else st.o_hi = o_k + d; }
  } else if (warp == 1) {
    const u32 d = gtd_tie_extent(o_k, s_k, -1);
    if (lane == 0) { if (d == 0xffffffffu) st.fail = 1; else st.o_lo = o_k - d; }
  }
  // everything outside topraw has a raw key below the smallest key in it, i.e. a logit <= that key's: scores <= e0
  const bool whole = m_top == n_keys && o_complete <= ord_f32(-INFINITY);
  const float e0 = whole ? -INFINITY : sigmoid_f32(unord_f32(st.min_o));
  if (warp >= 2)
  for (int i = tid - 64; i < m_top; i += RPP_EMIT_NT - 64) {   // (every slot has one owner: in place)
    const u64 rk2 = topraw[i];
    const float sc = sigmoid_f32(unord_f32((u32)(rk2 >> 32)));
    topraw[i] = sc > e0 ? make_key(sc, key_tie(rk2)) : 0ull;
  }
  __syncthreads();
  if (st.fail) return;
  const u32 o_lo = st.o_lo, o_hi = st.o_hi;
  // tie group: E = {o_lo <= logit <= o_hi}, n_gt = #{logit > o_hi}; rank the scored best keys meanwhile
  {
    int ngt = 0;
    for (int i = tid; i < n_keys; i += RPP_EMIT_NT) {
      const u64 k = sh->keys[i];
      const u32 o = (u32)(k >> 32);
      if (o > o_hi) ++ngt;
      else if (o >= o_lo) {
        const int slot = atomicAdd(&st.ne, 1);
        if (slot < RPP_GTD_TIES) eidx[slot] = key_tie(k);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ngt += __shfl_xor_sync(RPP_FULL_MASK, ngt, o);
    if (lane == 0 && ngt) atomicAdd(&st.ngt, ngt);
    // (8 lanes per key: the counting loop is the latency chain of this phase)
    for (int i0 = 0; i0 < m_top; i0 += RPP_EMIT_NT / 8) {
      const int i = i0 + (tid >> 3), sub = tid & 7;
      const u64 k = i < m_top ? topraw[i] : 0ull;
      int rank = 0;
      if (k != 0ull)
        for (int j = sub; j < m_top; j += 8) rank += topraw[j] > k;
      rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 1);
      rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 2);
      rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 4);
      if (k != 0ull && sub == 0) { top[rank] = k; atomicAdd(&st.nvalid, 1); }
    }
  }
  __syncthreads();
  const int ne = st.ne, nvalid = st.nvalid;
  const int slots = (int)P.k_lim - st.ngt;   // members of E inside the top k: its `slots` smallest flat indices
  // the tie group must lie inside what the keys cover, and fit; the best pairs must be complete
  if (ne > RPP_GTD_TIES || slots < 1 || slots > ne || o_lo < o_complete || nvalid < Mp) return;
  if (slots == ne) {
    if (tid == 0) st.idx_star = 0xffffffffu;
  } else {
    for (int i = tid; i < ne; i += RPP_EMIT_NT) {
      const u32 v = eidx[i];
      int rank = 0;
      for (int j = 0; j < ne; ++j) rank += eidx[j] < v;
      if (rank == slots - 1) st.idx_star = v;
    }
  }
  // anchors of the M best pairs (topraw is free again: [0, 1024) as u32 anchors, flags behind them)
  u32* anc = reinterpret_cast<u32*>(topraw);
  u32* isf = anc + 1024;
  const int got = Mp;
  if (tid < got) anc[tid] = key_tie(top[tid]) / (u32)G.C;
  __syncthreads();
  GTD_T();
  // ---- 3. one phase, no barrier inside: (a) thread i < M: box of pair i's anchor and "first pair of its anchor";
  //         (b) every class of every pair's anchor looked up in the logits -> candidate rows (filtered by (a) later)
  const u32 idx_star = st.idx_star;
  {
    float4 dl, an;
    u32 a_i = 0u;
    if (tid < got) {   // loads first, the dependent work after the look-ups were issued
      a_i = anc[tid];
      dl = lv_delta(G.src, b, a_i);
      an = G.anchors[a_i];
    }
    for (int e = tid; e < got * G.C; e += RPP_EMIT_NT) {
      const int i = e / G.C, c = e - i * G.C;
      const u32 a = anc[i];
      const float raw = lv_val(G.src, b, a, G.C, c);
      if (!(raw >= P.T_min)) continue;
      const u32 o = ord_f32(raw), flat = a * (u32)G.C + (u32)c;
      if (o > o_hi || (o >= o_lo && flat <= idx_star)) {
        const int slot = atomicAdd(&st.ncand, 1);
        if (slot < RPP_GTD_CAND) { cown[slot] = make_key(sigmoid_f32(raw), flat); cfirst[slot] = (u32)i; }
        else st.over = 1;
      }
    }
    for (int i0 = 0; i0 < got; i0 += RPP_EMIT_NT / 8) {   // "first pair of its anchor": 8 lanes per pair
      const int i = i0 + (tid >> 3), sub = tid & 7;
      int dup = 0;
      if (i < got) {
        const u32 a = anc[i];
        for (int j = sub; j < i; j += 8) dup |= anc[j] == a;
      }
      dup |= __shfl_xor_sync(RPP_FULL_MASK, dup, 1);
      dup |= __shfl_xor_sync(RPP_FULL_MASK, dup, 2);
      dup |= __shfl_xor_sync(RPP_FULL_MASK, dup, 4);
      // NonMaxSuppressionV5 never sees a row at or below the score threshold (A.2)
      if (i < got && sub == 0) isf[i] = (!dup && key_score(top[i]) > G.score_threshold) ? 1u : 0u;
    }
    if (tid < got) tbox[tid] = clip01(decode_box(dl, an, G.dp));
  }
  __syncthreads();
  if (st.over) return;   // the kernels that follow do the image
  GTD_T();
  // ---- 4. rank by counting on (anchor's best score desc, own pair key desc) = (row maximum desc, row index asc)
  const int nc_all = st.ncand;
  u32* cs = reinterpret_cast<u32*>(sh->keys);   // [nc_all] (the list is done) score bits of the candidate's anchor, 0 = not a row (pair is not its anchor's first)
  {
    int local = 0;
    for (int i = tid; i < nc_all; i += RPP_EMIT_NT) {
      const u32 f = cfirst[i];
      const u32 v = isf[f] ? (u32)(top[f] >> 32) : 0u;
      cs[i] = v;
      local += v != 0u;
    }
    if (local) atomicAdd(&st.nrows, local);
  }
  __syncthreads();
  const int nrows = st.nrows;
  const int valid = nrows < G.M ? nrows : G.M;
  if (tid == 0) { G.out_valid[b] = valid; P.emit_done[p] = 2; }
  for (int i0 = 0; i0 < nc_all; i0 += RPP_EMIT_NT / 8) {   // 8 lanes per candidate
    const int i = i0 + (tid >> 3), sub = tid & 7;
    const u32 sb = i < nc_all ? cs[i] : 0u;
    const u64 own = i < nc_all ? cown[i] : 0ull;
    int rank = 0;
    if (sb != 0u)
      for (int j = sub; j < nc_all; j += 8) {
        const u32 sj = cs[j];
        rank += (sj > sb) || (sj == sb && cown[j] > own);
      }
    rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 1);
    rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 2);
    rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 4);
    if (sb != 0u && sub == 0 && rank < G.M) {
      const size_t o = (size_t)b * G.M + rank;
      const u64 best = top[cfirst[i]];
      G.out_boxes[o] = tbox[cfirst[i]];
      G.out_scores[o] = key_score(best);
      G.out_classes[o] = (long long)(key_tie(best) % (u32)G.C);   // tf.argmax of the row: the anchor's best pair
    }
  }
  // padded selected index 0 -> boxes[0] (clipped), score -1, class -1 (:258-268)
  for (int i = valid + tid; i < G.M; i += RPP_EMIT_NT) {
    const size_t o = (size_t)b * G.M + i;
    G.out_boxes[o] = tbox[0];
    G.out_scores[o] = -1.0f;
    G.out_classes[o] = -1;
  }
  if (G.debug) {
    __syncthreads();
    GTD_T();
    if (tid == 0 && (b == 0 || b == 300))
      printf("gtd b=%d n_keys=%d m_top=%d nkbin=%d ne=%d rows=%d/%d prepare=%lld select=%lld ties+score=%lld cand=%lld out=%lld\n",
             b, n_keys, m_top, nkbin, ne, nrows, nc_all, tdbg[1] - tdbg[0], tdbg[2] - tdbg[1], tdbg[3] - tdbg[2],
             tdbg[4] - tdbg[3], tdbg[5] - tdbg[4]);
  }
#undef GTD_T
}


// ===============================================================================================================
// global_soft_kernel — NonMaxSuppressionV5 with soft_nms_sigma > 0 for ONE image per block, with the
// block's 16 warps working on one image instead of one warp popping a priority queue.
//
// The TF kernel pops the queue maximum, multiplies its score by the weights of the boxes selected since its last
// visit (newest first), and either selects it (score unchanged), drops it (score <= threshold) or pushes it back.
// Between two selections the pops are independent of each other — each only depends on the selected list — and the
// set of candidates visited in such a round is exactly a PREFIX of the queue order: candidates are visited while their
// (stale) key exceeds the largest re-scored key seen so far in the round; the first candidate whose score does not
// change is selected; if the stale keys fall below the largest re-scored key, that re-scored candidate is the next pop,
// finds nothing new to multiply and is selected.  Which candidates are visited in which round decides the grouping of
// the fp32 multiplications, so it is reproduced exactly; only the work inside a round runs in parallel:
//   1. the queue is ONE sorted ring of keys in shared memory (unvisited candidates are merged in from the sorted
//      stream in chunks, so the ring prefix is always the true queue prefix);
//   2. a batch = the first W keys; every candidate of the batch is re-scored by its own group of lanes against the
//      boxes selected since its last visit (speculatively: what lies beyond the round's cut is thrown away);
//   3. one warp finds the cut and the winner with a prefix maximum over the batch;
//   4. the visited prefix leaves the ring, the winner is appended to the selected list, the others re-enter the ring
//      at their new keys (parallel sorted insert: positions by search, one shifting pass from the tail).
// W = 16 (a warp per candidate); a round that outlives a batch switches to W = 128 (4 lanes per candidate): trained
// detectors put hundreds of overlapping candidates ahead of the next selection.
// ===============================================================================================================
#define RPP_GS_NT 512
#define RPP_GS_WMAX 128
#define RPP_GS_CHUNK 128      // stream keys merged into the ring at a time
#define RPP_GS_MAXK 8192

struct GlobalSoftParams {
  long k;                     // rows per image (<= RPP_GS_MAXK)
  int ring_cap;               // power of two >= k
  int box_cap;                // rows whose boxes are cached in shared memory (the rest is read through L2)
  float4* box_spill;          // [B][k] scratch: clipped boxes of rows >= box_cap (written and read by the block)
  const float4* anchors; DecodeParams dp;   // boxes are decoded when a row first enters the queue
  const u64* skey; const u64* dkey; const int* sd_cnt;
  float score_threshold, soft_scale, iou_threshold;
  int soft_ignores_iou;
  int M;
  int debug;                  // env RPP_GS_DEBUG: block 0 prints its phase cycle counts
  // outputs (global_out_kernel's contract)
  const u64* emit_key;        // [B][k]: row j -> anchor = tie / C
  Levels lv;                  // logits for the class lookup
  int C; long N;
  float4* out_boxes; float* out_scores; long long* out_classes; int* out_valid;
};

struct __align__(16) GlobalSoftShared {
  u64 stale[RPP_GS_WMAX];
  u64 fresh[RPP_GS_WMAX];     // 0 = dropped
  u64 ins[RPP_GS_WMAX];       // keys to insert, sorted descending
  u32 pos[RPP_GS_WMAX];       // ring elements greater than ins[i]
  u32 pos16[RPP_GS_NT / 32];  // narrow batches: rank of fresh[i] among the ring elements beyond the batch
  u32 omask[RPP_GS_NT / 32];  // narrow batches: batch slots whose box overlaps slot i's
  long long rdbg[8];          // debug (RPP_GS_DEBUG): cycle counts inside the re-scoring of batch slot 1
  int head, cnt;              // ring
  int s_pos, s_n;             // sorted stream of unvisited first-occurrence rows
  int nsel;
  int cut, winner, n_ins, finished;
  u64 s_next;                 // next unmerged stream key (0 when the stream is exhausted)
  u64 exp_tab[32];            // shared-memory copy of c_exp2f_tab
  // dynamic: u64 ring[ring_cap]; float4 kbox[M]; float karea[M]; u64 sel[M]; float4 boxes[box_cap]; u16 begin[k]
};

__host__ __device__ static inline size_t global_soft_smem(long k, int ring_cap, int box_cap, int M) {
  return ((sizeof(GlobalSoftShared) + 15) & ~(size_t)15) + (size_t)ring_cap * 8 + (size_t)M * 16 +
         (size_t)box_cap * 16 + ((((size_t)M * 8) + 15) & ~(size_t)15) + ((((size_t)M * 4) + 15) & ~(size_t)15) +
         (((size_t)k * 2 + 15) & ~(size_t)15);
}

// pos[i] = number of ring elements greater than ins[i] (binary search).  Called by the thread that owns key i.
__device__ __forceinline__ u32 gs_ring_rank(const u64* ring, int mask, int head, int cnt, u64 key) {
  int lo = 0, hi = cnt;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ring[(head + mid) & mask] > key) lo = mid + 1; else hi = mid;
  }
  return (u32)lo;
}

// The same by a whole warp (all 32 lanes call with the same key): 32-ary search, one probe per lane and round.
__device__ __forceinline__ u32 gs_ring_rank_warp(const u64* ring, int mask, int head, int cnt, u64 key) {
  const int lane = threadIdx.x & 31;
  int lo = 0, hi = cnt;   // the answer is in [lo, hi]
  while (hi - lo > 32) {
    const int stride = (hi - lo + 31) >> 5;
    const int p = lo + lane * stride;
    const u32 gt = __ballot_sync(RPP_FULL_MASK, p < hi && ring[(head + p) & mask] > key);
    const int nb = __popc(gt);           // probes 0 .. nb-1 are greater (the ring is sorted descending)
    const int nlo = nb > 0 ? lo + (nb - 1) * stride + 1 : lo;
    const int nhi = lo + nb * stride < hi ? lo + nb * stride : hi;
    lo = nlo; hi = nhi;
  }
  const u32 gt = __ballot_sync(RPP_FULL_MASK, lo + lane < hi && ring[(head + lo + lane) & mask] > key);
  return (u32)(lo + __popc(gt));
}

// Moves the ring elements aside and writes the n_ins keys of sh->ins (sorted descending) at pos[i] + i; pos[] is
// already computed and visible.  All threads call.  Two barriers per 512 moved elements.
__device__ void gs_ring_shift_insert(GlobalSoftShared* sh, u64* ring, int mask) {
  const int tid = threadIdx.x;
  const int m = sh->n_ins;
  if (m == 0) return;
  const int head = sh->head, cnt = sh->cnt;
  // from the tail, in slabs of one element per thread: element p moves to p + #{i : pos[i] <= p}
  const int first = (int)sh->pos[0];
  // nothing to move (every new key lands behind the ring): the barrier the slab loop would have provided — every
  // thread has read the ring state above before thread 0 updates the count below
  if (cnt <= first) __syncthreads();
  for (int hi = cnt; hi > first; hi -= RPP_GS_NT) {
    const int p = hi - 1 - tid;
    u64 v = 0ull;
    int g = 0;
    if (p >= first) {
      v = ring[(head + p) & mask];
      int lo2 = 0, hi2 = m;   // upper bound of p in pos[] (non-decreasing)
      while (lo2 < hi2) {
        const int mid = (lo2 + hi2) >> 1;
        if ((int)sh->pos[mid] <= p) lo2 = mid + 1; else hi2 = mid;
      }
      g = lo2;
    }
    __syncthreads();
    if (p >= first && g > 0) ring[(head + p + g) & mask] = v;
    if (hi - RPP_GS_NT > first) __syncthreads();   // (the last slab shares its barrier with the key writes below)
  }
  // every source position has been read: the new keys go to their slots
  if (tid < m) ring[(head + (int)sh->pos[tid] + tid) & mask] = sh->ins[tid];
  if (tid == 0) sh->cnt = cnt + m;
  __syncthreads();
}

// Inserts the n_ins keys of sh->ins (sorted descending) into the sorted ring.  All threads call.
__device__ void gs_ring_insert(GlobalSoftShared* sh, u64* ring, int mask) {
  const int tid = threadIdx.x;
  const int m = sh->n_ins;
  if (m == 0) return;
  if (tid < m) sh->pos[tid] = gs_ring_rank(ring, mask, sh->head, sh->cnt, sh->ins[tid]);
  __syncthreads();
  gs_ring_shift_insert(sh, ring, mask);
}

// Box of filtered row j, clipped to [0,1] (every mode but CombinedNMS clips before NMS, :275): decoded from the row's
// anchor when the row enters the queue, cached in shared memory (rows < box_cap) or in the block's spill area.
__device__ __forceinline__ float4 gs_decode_row(const GlobalSoftParams& P, int b, u32 j) {
  const u64 key = P.emit_key[(size_t)b * P.k + j];
  if (key == 0ull) return make_float4(0.f, 0.f, 0.f, 0.f);   // (no such row: fewer than k finite logits)
  const u32 a = key_tie(key) / (u32)P.C;
  return clip01(decode_box(lv_delta(P.lv, b, a), P.anchors[a], P.dp));
}
__device__ __forceinline__ void gs_store_box(const GlobalSoftParams& P, float4* boxes, float4* spill, u32 j, float4 v) {
  if ((int)j < P.box_cap) boxes[j] = v; else __stcg(&spill[j], v);
}
__device__ __forceinline__ float4 gs_load_box(const GlobalSoftParams& P, const float4* boxes, const float4* spill, u32 j) {
  return (int)j < P.box_cap ? boxes[j] : __ldcg(&spill[j]);
}

template <int LPC>
__device__ __forceinline__ void gs_refresh(const GlobalSoftParams& P, GlobalSoftShared* sh, const u64* ring, int mask,
                                           const float4* kbox, const float* karea, const float4* boxes,
                                           const unsigned short* begin, const float4* gboxes, int batch) {
  const int tid = threadIdx.x;
  const int ci = tid / LPC, gl = tid % LPC;
  if (ci >= batch) return;
  const u32 gmask = LPC == 32 ? RPP_FULL_MASK : (((1u << LPC) - 1u) << ((tid & 31) / LPC * LPC));
  const u64 st = ring[(sh->head + ci) & mask];
  const u32 j = key_tie(st);
  float score = key_score(st);
  const int bg = begin[j];
  const int nsel = sh->nsel;
  float4 box = gs_load_box(P, boxes, gboxes, j);
  float area;
  {
    const float4 cb = canon_box(box, area);
    if (area > 0.0f) box = cb; else { box = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY); area = 0.0f; }
  }
  const float thr = P.score_threshold;
  bool dropped = false;
  for (int top = nsel - 1; top >= bg && !dropped; top -= LPC) {
    const int jj = top - gl;
    float w = 1.0f;
    if (jj >= bg) {
      const float sim = iou_val(box, area, kbox[jj], karea[jj]);
      // sim == 0 (no overlap, the common case): expf(scale * 0 * 0) = expf(0) = 1 exactly
      if (sim != 0.0f) w = expf_glibc_tab(__fmul_rn(__fmul_rn(P.soft_scale, sim), sim), sh->exp_tab);
      if (!P.soft_ignores_iou && sim > P.iou_threshold) w = 0.0f;
    }
    // multiply in the TF kernel's order: newest selected first = ascending lane of the group
    u32 nz = __ballot_sync(gmask, w != 1.0f) & gmask;
    while (nz) {
      const int t = __ffs(nz) - 1;
      nz &= nz - 1u;
      score = __fmul_rn(score, __shfl_sync(gmask, w, t));
      if (score <= thr) { dropped = true; break; }
    }
  }
  if (gl == 0) {
    sh->stale[ci] = st;
    sh->fresh[ci] = (dropped || !(score > thr)) ? 0ull : make_key(score, j);
  }
}

// Narrow batches: a warp per candidate.  Most (candidate, selected box) pairs do not overlap, and the expensive part of
// a visit is the weight (IEEE division + binary64 expf) of the few that do.  Pass 1 only flags the overlapping pairs
// (four strips of 32 selected boxes, one ballot each); pass 2 hands the flagged pairs, in the TF kernel's
// multiplication order (newest selected first), one to a lane and evaluates all their weights at once; the
// multiplications then follow in that order.  The warp finally ranks the re-scored key in the ring beyond the batch,
// so that the insert position is known before the cut is.
__device__ __forceinline__ void gs_refresh_warp(const GlobalSoftParams& P, GlobalSoftShared* sh, const u64* ring,
                                                int mask, const float4* kbox, const float* karea, const float4* boxes,
                                                const unsigned short* begin, const float4* gboxes, int batch, int cnt) {
  const int ci = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ci >= batch) return;
  const bool dbgw = P.debug && blockIdx.x == 0 && ci == 1 && lane == 0;
  long long tq = dbgw ? clock64() : 0;
#define RDBG(slot) do { if (dbgw) { const long long t__ = clock64(); sh->rdbg[slot] += t__ - tq; tq = t__; } } while (0)
  const int head = sh->head;
  const u64 st = ring[(head + ci) & mask];
  const u32 j = key_tie(st);
  float score = key_score(st);
  const int bg = begin[j];
  const int nsel = sh->nsel;
  float4 box = gs_load_box(P, boxes, gboxes, j);
  float area;
  {
    const float4 cb = canon_box(box, area);
    if (area > 0.0f) box = cb; else { box = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY); area = 0.0f; }
  }
  const float thr = P.score_threshold;
  bool dropped = false;
  RDBG(0);
  for (int top = nsel - 1; top >= bg && !dropped; top -= 128) {
    u32 m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = top - 32 * u - lane;
      bool ov = false;
      if (jj >= bg) {
        const float4 kb = kbox[jj];
        ov = fminf(box.z, kb.z) > fmaxf(box.x, kb.x) && fminf(box.w, kb.w) > fmaxf(box.y, kb.y);
      }
      m[u] = __ballot_sync(RPP_FULL_MASK, ov);
    }
    const int c0 = __popc(m[0]), c1 = __popc(m[1]), c2 = __popc(m[2]), c3 = __popc(m[3]);
    const int total = c0 + c1 + c2 + c3;
    RDBG(1);
    for (int base = 0; base < total && !dropped; base += 32) {
      const int n = base + lane;
      float w = 1.0f;
      if (n < total) {
        int r = n, u = 0;
        u32 ms = m[0];
        if (r >= c0) { r -= c0; u = 1; ms = m[1]; if (r >= c1) { r -= c1; u = 2; ms = m[2]; if (r >= c2) { r -= c2; u = 3; ms = m[3]; } } }
        const int jj = top - 32 * u - (int)__fns(ms, 0u, r + 1);
        const float sim = iou_val(box, area, kbox[jj], karea[jj]);
        w = expf_glibc_tab(__fmul_rn(__fmul_rn(P.soft_scale, sim), sim), sh->exp_tab);   // sim == 0: exactly 1
        if (!P.soft_ignores_iou && sim > P.iou_threshold) w = 0.0f;
      }
      u32 nz = __ballot_sync(RPP_FULL_MASK, w != 1.0f);
      while (nz) {
        const int t = __ffs(nz) - 1;
        nz &= nz - 1u;
        score = __fmul_rn(score, __shfl_sync(RPP_FULL_MASK, w, t));
        if (score <= thr) { dropped = true; break; }
      }
    }
  }
  RDBG(2);
  const u64 fk = (dropped || !(score > thr)) ? 0ull : make_key(score, j);
  u32 r16 = 0u;
  if (fk != 0ull && fk != st) r16 = gs_ring_rank_warp(ring, mask, (head + batch) & mask, cnt - batch, fk);
  RDBG(3);
  // which other slots of the batch overlap this one (their weight against it can differ from 1)
  bool ov = false;
  if (lane < batch && lane != ci) {
    float4 ob = gs_load_box(P, boxes, gboxes, key_tie(ring[(head + lane) & mask]));
    float oa;
    const float4 oc = canon_box(ob, oa);
    ov = oa > 0.0f && fminf(box.z, oc.z) > fmaxf(box.x, oc.x) && fminf(box.w, oc.w) > fmaxf(box.y, oc.y);
  }
  const u32 om = __ballot_sync(RPP_FULL_MASK, ov);
  if (lane == 0) { sh->stale[ci] = st; sh->fresh[ci] = fk; sh->pos16[ci] = r16; sh->omask[ci] = om; }
  RDBG(4);
#undef RDBG
}

#define GS_T(slot) do { if (P.debug && blockIdx.x == 0 && threadIdx.x == 0) { const long long t__ = clock64(); dbg[slot] += t__ - tlast; tlast = t__; } } while (0)
__global__ void __launch_bounds__(RPP_GS_NT) global_soft_kernel(GlobalSoftParams P) {
  pdl_enter();
  long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  int n_batches = 0, n_wide = 0, n_merge = 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GlobalSoftShared* sh = reinterpret_cast<GlobalSoftShared*>(smem_raw);
  // explicit 16-byte aligned offsets (see global_soft_smem)
  size_t off = (sizeof(GlobalSoftShared) + 15) & ~(size_t)15;
  u64* ring = reinterpret_cast<u64*>(smem_raw + off);            off += (size_t)P.ring_cap * 8;
  float4* kbox = reinterpret_cast<float4*>(smem_raw + off);      off += (size_t)P.M * 16;
  float4* boxes = reinterpret_cast<float4*>(smem_raw + off);     off += (size_t)P.box_cap * 16;
  u64* sel = reinterpret_cast<u64*>(smem_raw + off);             off += (((size_t)P.M * 8) + 15) & ~(size_t)15;
  float* karea = reinterpret_cast<float*>(smem_raw + off);       off += (((size_t)P.M * 4) + 15) & ~(size_t)15;
  unsigned short* begin = reinterpret_cast<unsigned short*>(smem_raw + off);
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.x;
  const int mask = P.ring_cap - 1;
  float4* gboxes = P.box_spill + (size_t)b * P.k;
  const u64* skey = P.skey + (size_t)b * P.k;

  // ---- prologue: boxes, visit state, the duplicate rows sorted into the ring -----------------------------------
  for (long j = tid; j < P.k; j += RPP_GS_NT) begin[j] = 0;
  const int nd = P.sd_cnt[2 * b + 1];
  for (int i = tid; i < nd; i += RPP_GS_NT) {   // the duplicate rows start in the queue
    const u32 j = key_tie(P.dkey[(size_t)b * P.k + i]);
    gs_store_box(P, boxes, gboxes, j, gs_decode_row(P, b, j));
  }
  int P2 = 2;
  while (P2 < nd) P2 <<= 1;
  for (int i = tid; i < P2 && nd > 0; i += RPP_GS_NT) ring[i] = i < nd ? P.dkey[(size_t)b * P.k + i] : 0ull;
  if (tid < 32) sh->exp_tab[tid] = c_exp2f_tab[tid];
  if (tid < 8) sh->rdbg[tid] = 0;
  if (tid == 0) {
    const int ns = P.sd_cnt[2 * b];
    sh->head = 0; sh->cnt = nd;
    sh->s_pos = 0; sh->s_n = ns;
    sh->s_next = ns > 0 ? skey[0] : 0ull;
    sh->nsel = 0; sh->finished = 0;
  }
  __syncthreads();
  if (nd > 1) bitonic_sort_desc<RPP_GS_NT>(ring, P2);

  GS_T(0);
  int wide = 0;
  for (;;) {
    const int W = wide ? RPP_GS_WMAX : RPP_GS_NT / 32;
    // ---- 1. the ring prefix must be the queue prefix: merge stream chunks until ring[W-1] beats the stream head ---
    for (;;) {
      const int s_pos = sh->s_pos, s_n = sh->s_n, cnt = sh->cnt;
      if (s_pos >= s_n) break;
      if (cnt >= W && ring[(sh->head + W - 1) & mask] > sh->s_next) break;
      const int m = s_n - s_pos < RPP_GS_CHUNK ? s_n - s_pos : RPP_GS_CHUNK;
      __syncthreads();
      if (tid < m) {
        const u64 key = skey[s_pos + tid];
        sh->ins[tid] = key;
        gs_store_box(P, boxes, gboxes, key_tie(key), gs_decode_row(P, b, key_tie(key)));
      }
      if (tid == 0) {
        sh->n_ins = m; sh->s_pos = s_pos + m;
        sh->s_next = s_pos + m < s_n ? skey[s_pos + m] : 0ull;
      }
      __syncthreads();
      gs_ring_insert(sh, ring, mask);
      ++n_merge;
    }
    GS_T(1);
    const int cnt = sh->cnt;
    if (cnt == 0 || sh->nsel >= P.M) break;
    ++n_batches; n_wide += wide;
    const int batch = cnt < W ? cnt : W;
    const bool all = batch == cnt && sh->s_pos >= sh->s_n;   // the batch is everything that is left
    // ---- 2. speculative re-scoring --------------------------------------------------------------------------------
    if (wide) gs_refresh<4>(P, sh, ring, mask, kbox, karea, boxes, begin, gboxes, batch);
    else gs_refresh_warp(P, sh, ring, mask, kbox, karea, boxes, begin, gboxes, batch, cnt);
    __syncthreads();
    GS_T(2);
    if (!wide) {
      // ---- 3 + 4 (narrow batches, <= 16 candidates): warp 0 finds the cut and the winner, commits the visited
      // prefix, sorts the keys that re-enter the ring and finds their positions, all without a block barrier ---------
      if (tid < 32) {
        // Lane = batch slot.  The warp replays the TF kernel's pops over the batch as far as the re-scored keys of
        // pass 2 stay valid: a candidate's re-scoring (against the boxes selected before the batch) is still its
        // re-scoring at the time of its pop when no box selected INSIDE the batch since then overlaps it (weight
        // exactly 1) — so one batch usually carries several selections; it ends at the first pop that would need
        // a weight the batch does not have, at the end of the batch, or with the selected list full.
        const int cnt0 = cnt;
        const int nsel0 = sh->nsel, head0 = sh->head;   // (read by every lane before lane 0 updates them below)
        const u64 st = lane < batch ? sh->stale[lane] : 0ull;
        const u64 fr = lane < batch ? sh->fresh[lane] : 0ull;
        const u32 om = lane < batch ? sh->omask[lane] : 0u;
        const u64 st_last = __shfl_sync(RPP_FULL_MASK, st, batch - 1);
        int status = 0;          // 0 not visited, 1 visited and back in the queue, 2 selected, 3 dropped
        int vbeg = 0, ord = -1;  // selected.size() at the visit; selection order inside the batch
        u32 sbv = 0u;            // boxes of the batch selected before this slot's visit
        u32 Sb = 0u;             // boxes of the batch selected so far
        int nsel_cur = nsel0, nselb = 0, next_unv = 0;
        u64 vmax = 0ull;         // queue maximum among the slots that are back in the queue
        int vslot = -1;
        for (;;) {
          if (nsel_cur >= P.M) break;
          const u64 ukey = next_unv < batch ? __shfl_sync(RPP_FULL_MASK, st, next_unv) : 0ull;
          bool pop_v;
          if (next_unv >= batch) {
            // the next never-visited candidate lies beyond the batch (its key is below the batch's last)
            if (vmax == 0ull) break;
            if (!all && vmax < st_last) break;
            pop_v = true;
          } else {
            pop_v = vmax > ukey;
          }
          if (pop_v) {
            const u32 om_y = __shfl_sync(RPP_FULL_MASK, om, vslot), sbv_y = __shfl_sync(RPP_FULL_MASK, sbv, vslot);
            if (om_y & Sb & ~sbv_y) break;        // a box selected since its visit overlaps it: a real re-scoring
            if (lane == vslot) { status = 2; ord = nselb; }
            Sb |= 1u << vslot; ++nselb; ++nsel_cur;
            u64 mx = status == 1 ? fr : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const u64 other = __shfl_xor_sync(RPP_FULL_MASK, mx, o);
              mx = other > mx ? other : mx;
            }
            vmax = mx;
            const u32 who = __ballot_sync(RPP_FULL_MASK, status == 1 && fr == mx);
            vslot = (mx != 0ull) ? __ffs(who) - 1 : -1;
          } else {
            const int x = next_unv;
            if (__shfl_sync(RPP_FULL_MASK, om, x) & Sb) break;   // overlaps a box selected inside the batch
            const u64 fr_x = __shfl_sync(RPP_FULL_MASK, fr, x);
            if (fr_x == ukey) {                    // score unchanged: selected
              if (lane == x) { status = 2; ord = nselb; }
              Sb |= 1u << x; ++nselb; ++nsel_cur;
            } else if (fr_x == 0ull) {             // fell to the threshold
              if (lane == x) status = 3;
            } else {                               // back into the queue, re-scored
              if (lane == x) { status = 1; vbeg = nsel_cur; sbv = Sb; }
              if (fr_x > vmax) { vmax = fr_x; vslot = x; }
            }
            ++next_unv;
          }
        }
        const int cut = next_unv;
        // commit
        u64 mine = 0ull;
        if (lane < batch && status != 0) {
          const u32 j = key_tie(st);
          if (status == 2) {
            float4 box = gs_load_box(P, boxes, gboxes, j);
            float area;
            const float4 cb = canon_box(box, area);
            kbox[nsel0 + ord] = area > 0.0f ? cb : make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
            karea[nsel0 + ord] = area > 0.0f ? area : 0.0f;
            sel[nsel0 + ord] = fr;
          } else if (status == 1) {
            begin[j] = (unsigned short)vbeg;      // suppress_begin_index = selected.size() at the visit
            mine = fr;
          }
        }
        // The keys that re-enter the ring: rank among themselves (sorted descending; keys are unique) and position
        // among the ring elements that stay = rank beyond the batch (found by the candidate's warp during the
        // re-scoring) + the unvisited rest of the batch.
        const u64 pk = lane < batch ? (lane < cut ? mine : st) : 0ull;
        int rank = 0, stay = 0;
        for (int l = 0; l < batch; ++l) {
          const u64 o = __shfl_sync(RPP_FULL_MASK, pk, l);
          const bool gt = o > mine;
          rank += gt && l < cut;
          stay += gt && l >= cut;
        }
        const int m = __popc(__ballot_sync(RPP_FULL_MASK, mine != 0ull));
        const int head2 = (head0 + cut) & mask, cnt2 = cnt0 - cut;
        if (mine != 0ull) {
          sh->ins[rank] = mine;
          sh->pos[rank] = sh->pos16[lane] + (u32)stay;
        }
        __syncwarp();
        if (lane == 0) {
          sh->n_ins = m; sh->head = head2; sh->cnt = cnt2;
          sh->cut = cut; sh->winner = nselb > 0 ? 0 : -1;
          sh->nsel = nsel_cur;
        }
      }
      __syncthreads();
      GS_T(3);
      gs_ring_shift_insert(sh, ring, mask);
      if (sh->n_ins == 0) __syncthreads();    // (shift_insert returns at once when nothing re-enters)
      GS_T(4);
      const int winner = sh->winner;
      if (sh->finished) break;
      wide = (winner < 0) ? 1 : 0;
      continue;
    }
    // ---- 3. cut and winner (warp 0; lane l owns batch slots 4l .. 4l+3) ---------------------------------------------
    if (tid < 32) {
      u64 st4[4], fr4[4], ex4[4];
      u64 run = 0ull;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = 4 * lane + u;
        st4[u] = i < batch ? sh->stale[i] : 0ull;
        fr4[u] = i < batch ? sh->fresh[i] : 0ull;
        ex4[u] = run;                       // maximum of the lane's earlier slots
        run = fr4[u] > run ? fr4[u] : run;
      }
      u64 incl = run;                       // inclusive prefix maximum over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 up = __shfl_up_sync(RPP_FULL_MASK, incl, o);
        if (lane >= o && up > incl) incl = up;
      }
      u64 excl = __shfl_up_sync(RPP_FULL_MASK, incl, 1);
      if (lane == 0) excl = 0ull;
      int stop = 1 << 20, stop_kind = 0;    // first slot where the round ends: 1 = unchanged candidate, 2 = stale < max
#pragma unroll
      for (int u = 3; u >= 0; --u) {
        const int i = 4 * lane + u;
        if (i >= batch) continue;
        const u64 before = ex4[u] > excl ? ex4[u] : excl;   // largest re-scored key among slots < i
        if (before > st4[u]) { stop = i; stop_kind = 2; }
        else if (fr4[u] == st4[u]) { stop = i; stop_kind = 1; }
      }
      int best = stop;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(RPP_FULL_MASK, best, o);
        best = other < best ? other : best;
      }
      const u32 owner = __ballot_sync(RPP_FULL_MASK, stop == best && best < (1 << 20));
      const int kind = owner ? __shfl_sync(RPP_FULL_MASK, stop_kind, __ffs(owner) - 1) : 0;
      int cut = batch, winner = -1;
      if (kind == 1) { cut = best + 1; winner = best; }
      else if (kind == 2 || all) {
        // the largest re-scored key among the visited slots [0, cut) is the next pop and is selected
        cut = kind == 2 ? best : batch;
        u64 mx = 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (4 * lane + u < cut && fr4[u] > mx) mx = fr4[u];
        u64 wmx = mx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const u64 other = __shfl_xor_sync(RPP_FULL_MASK, wmx, o);
          wmx = other > wmx ? other : wmx;
        }
        if (wmx != 0ull) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (4 * lane + u < cut && fr4[u] == wmx) winner = 4 * lane + u;
          const u32 who = __ballot_sync(RPP_FULL_MASK, winner >= 0);
          winner = __shfl_sync(RPP_FULL_MASK, winner, __ffs(who) - 1);
        } else if (all) {
          if (lane == 0) sh->finished = 1;   // everything that was left fell to the threshold
        }
      }
      if (lane == 0) { sh->cut = cut; sh->winner = winner; }
    }
    __syncthreads();
    // ---- 4. commit ----------------------------------------------------------------------------------------------
    const int cut = sh->cut, winner = sh->winner;
    const int nsel = sh->nsel;
    u64 mine = 0ull;
    if (tid < cut) {
      const u64 fr = sh->fresh[tid];
      const u32 j = key_tie(sh->stale[tid]);
      begin[j] = (unsigned short)nsel;          // suppress_begin_index = selected.size() at the visit
      if (tid == winner) {
        float4 box = gs_load_box(P, boxes, gboxes, j);
        float area;
        const float4 cb = canon_box(box, area);
        kbox[nsel] = area > 0.0f ? cb : make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
        karea[nsel] = area > 0.0f ? area : 0.0f;
        sel[nsel] = fr;
      } else {
        mine = fr;
      }
    }
    // keys that re-enter the ring, sorted descending (rank by counting; keys are unique)
    if (tid < RPP_GS_WMAX) sh->pos[tid] = 0u;
    __syncthreads();
    if (tid < cut && mine != 0ull) {
      int rank = 0;
      for (int i = 0; i < cut; ++i) {
        const u64 o = (i == winner) ? 0ull : sh->fresh[i];
        rank += o > mine;
      }
      sh->pos[rank] = 1u;                       // (marks the slot; pos[] is recomputed by the insert)
      sh->ins[rank] = mine;
    }
    __syncthreads();
    if (tid == 0) {
      int m = 0;
      while (m < cut && sh->pos[m]) ++m;        // ranks are dense: 0 .. m-1
      sh->n_ins = m;
      sh->head = (sh->head + cut) & mask;
      sh->cnt = cnt - cut;
      if (winner >= 0) sh->nsel = nsel + 1;
    }
    __syncthreads();
    gs_ring_insert(sh, ring, mask);
    if (sh->finished) break;
    // a round that outlived its batch has many overlapping candidates ahead of the next selection: go wide
    wide = (winner < 0) ? 1 : 0;
  }
  __syncthreads();
  GS_T(5);
  // ---- epilogue: GenerateDetections._global_nms outputs (:258-268) -----------------------------------------------
  const int valid = sh->nsel;
  if (tid == 0) P.out_valid[b] = valid;
  // one warp per selected row: tf.argmax over the row's scores = first class whose SCORE equals the row maximum.
  // The maximum logit and its first index are found on raw logits; a lower logit can only round to the same score when
  // it is close to the maximum (or when the scores saturate), so the sigmoid is evaluated for those few only.
  for (int i = tid >> 5; i < P.M; i += RPP_GS_NT / 32) {
    const size_t o = (size_t)b * P.M + i;
    if (i < valid) {
      const u64 key = sel[i];
      const u32 j = key_tie(key);
      const u32 a = key_tie(P.emit_key[(size_t)b * P.k + j]) / (u32)P.C;
      // one pass: maximum logit and the lowest class index that holds it
      float best = -INFINITY;
      int cls = 0x7fffffff;
      for (int c = lane; c < P.C; c += 32) {
        const float raw = lv_val(P.lv, b, a, P.C, c);
        if (raw > best) { best = raw; cls = c; }
      }
#pragma unroll
      for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const float ob = __shfl_xor_sync(RPP_FULL_MASK, best, ofs);
        const int oc = __shfl_xor_sync(RPP_FULL_MASK, cls, ofs);
        if (ob > best || (ob == best && oc < cls)) { best = ob; cls = oc; }
      }
      // a lower logit in an earlier class that rounds to the same score wins the argmax: only possible close to the
      // maximum or where the sigmoid saturates
      bool near = false;
      for (int c = lane; c < cls; c += 32) {
        const float raw = lv_val(P.lv, b, a, P.C, c);
        near = near || raw > best - 1.0f || best > 15.0f || best < -80.0f;
      }
      if (__any_sync(RPP_FULL_MASK, near)) {
        const float s_best = sigmoid_f32(best);
        int c2 = 0x7fffffff;
        for (int c = lane; c < cls; c += 32)
          if (sigmoid_f32(lv_val(P.lv, b, a, P.C, c)) == s_best && c < c2) c2 = c;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) c2 = min(c2, __shfl_xor_sync(RPP_FULL_MASK, c2, ofs));
        if (c2 < cls) cls = c2;
      }
      if (lane == 0) {
        P.out_boxes[o] = gs_load_box(P, boxes, gboxes, j);
        P.out_scores[o] = key_score(key);
        P.out_classes[o] = cls;
      }
    } else if (lane == 0) {
      P.out_boxes[o] = gs_decode_row(P, b, 0u);   // padded selected index 0 -> boxes[0] (clipped), score -1, class -1
      P.out_scores[o] = -1.0f;
      P.out_classes[o] = -1;
    }
  }
  if (P.debug && blockIdx.x == 0) {
    __syncthreads();
    GS_T(6);
    if (threadIdx.x == 0)
      printf("global_soft: batches %d (wide %d) merges %d nsel %d | cycles prologue %lld stage1 %lld refresh %lld decide %lld shift %lld loop-exit %lld epilogue %lld\n",
             n_batches, n_wide, n_merge, sh->nsel, dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[5], dbg[6]);
    if (threadIdx.x == 0)
      printf("global_soft refresh(slot 1): setup %lld flags %lld pairs+product %lld rank %lld overlap-mask %lld\n",
             sh->rdbg[0], sh->rdbg[1], sh->rdbg[2], sh->rdbg[3], sh->rdbg[4]);
  }
}
