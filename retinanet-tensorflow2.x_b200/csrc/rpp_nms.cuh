// rpp_nms.cuh — K3: per-problem lazy selection + NMS consumers (hard, padded, soft, top-k emission), probe and bound kernels.
// Part of the retinapost kernel set; included by rpp_kernels.cuh (one translation unit: rpp_api.cu).
#pragma once
#include "rpp_kernels.cuh"

// ===============================================================================================================
// K3  per-(image, class) problem kernel: lazy exact selection + greedy hard NMS
//     (CombinedNMS per-class stage;  PerClassHardNMS = NonMaxSuppressionV5 hard, A.2).
// ===============================================================================================================
#ifndef RPP_NMS_NT
#define RPP_NMS_NT 128
#endif
#define RPP_CHUNK_CAP 1024   // merge kernel
#define RPP_NMS_CHUNK 512
#define RPP_LIST_SMEM 1536

struct ColProblemParams {
  // source: columns of a [B, N, C] tensor
  Levels lv;               // x: logits (is_logit = 1) or scores (is_logit = 0); d: box deltas (fused path)
  int is_logit;
  long N;                  // rows per image
  int C;
  // boxes: decoded on demand (deltas + anchors) or gathered from a dense [B, N, q, 4] tensor
  const float4* anchors;   // [N]
  const float4* boxes;     // dense boxes (stage-wise) or nullptr
  int q;
  // rows that are not rows of the delta tensor (the global filter without the row gather, rpp_global.cuh): row j is
  // anchor tie(row_keys[b][j]) / C_src of the deltas in dlv
  const u64* row_keys; long k_rows; int C_src;
  Levels dlv;              // dlv.L > 0 without row_keys: the deltas of row r are dlv's row r (x is a derived column)
  DecodeParams dp;
  int clip_before;         // clip boxes to [0,1] before IoU (every mode but CombinedNMS; B6)
  float iou_threshold;
  float score_threshold;
  float T_min;             // raw pre-image of the score threshold (candidates have raw >= T_min)
  int M_lim;               // max kept per problem (also sizes the kept arrays in shared memory)
  // Two-pass scheme of the per-class modes (DESIGN.md "cross-class bound"): pass 1 (probe) keeps at most M_cap = m1
  // boxes per class and records `bound` = score of its last kept box (-inf when the class is exhausted); a tiny
  // kernel turns the probes of an image into stop_L = a lower bound of the image's M-th best final score; pass 2
  // re-runs only the classes whose bound >= stop_L, stopping at the first candidate below stop_L.
  int pass;                // 0 single pass, 1 probe, 2 finish
  int argmax;              // finish pass of the hard modes: argmax-iterate consumer (dynamic shared memory holds
                           // RPP_LIST_SMEM candidate boxes behind the kept arrays)
  int M_cap;               // kept limit of this pass
  int want0;               // size of the first chunk
  float* bound;            // [P]
  const float* stop_L;     // [B] or nullptr
  // finish pass: the bound kernel lists the problems that still need work (usually 1-3 % of them) and a few
  // persistent blocks pop them, instead of launching P blocks of which almost all exit at once
  const u32* work_items;   // [P] or nullptr (one block per problem)
  u32* work_ctl;           // [0] = number of items, [1] = pop cursor
  long k_lim;              // max candidates consumed (pre_nms_top_k after clamping; N when unfiltered)
  int M;                   // stride of the sel_* arrays
  // candidate lists
  const float* T;          // [P] thresholds used by the collect pass
  u32* cand_count;         // [P]; bit 31 = the list was already converted to keys in place (long lists)
  uint2* cand;             // [P][CAP]
  int CAP;
  int force_scan;          // debug: ignore the lists, use the exact column scan only
  u32* scan_count;         // handle-wide counter: problems that went on to an exact scan of their column
  // outputs per problem
  u64* sel_key;            // [P][M]  (final score bits | ~row index)
  float4* sel_box;         // [P][M]  kept boxes as they leave NMS (clipped iff clip_before)
  int* sel_cnt;            // [P]
  // soft NMS (NonMaxSuppressionV5 with soft_nms_sigma > 0)
  float soft_scale;        // -0.5 / soft_nms_sigma (soft_nms_sigma = config sigma / 2)
  int soft_ignores_iou;    // TF >= 2.3 weight form
  int tie_is_rank;         // NMS index of a candidate = its rank in the filtered list (per-class top-k ran first)
  u64* r_key;              // [P][r_cap] spill of the re-scored queue beyond shared memory
  uint2* r_meta;
  float4* r_box;
  long r_cap;
  // top-k emission (FilterTopKDetections): sorted keys of the k_lim best rows
  u64* emit_key;           // [P][k_lim]
  int* emit_done;          // [P] 1 = emit_sort_kernel already wrote this problem's keys; 2 = global_top_direct_kernel
                           // already wrote the image's detections (nothing downstream has work for it)
  int emit_direct;         // emit_done was initialised by global_top_direct_kernel
  long n_loop;             // > 0: a few persistent blocks walk problems [0, n_loop) (almost all of them are already done:
                           // the kernels behind global_top_direct_kernel); 0: one block per problem
  // tf.image.non_max_suppression_padded semantics (the TPU branches, postprocessing_ops.py:288-432; consumer
  // RPP_CONSUME_PADDED): 1 = _tpu_global_hard_nms (score filter inside), 2 = _tpu_per_class_hard_nms (every row is
  // a candidate; score_threshold / T_min of this struct are -inf and stop_score holds the config threshold)
  int padded;
  float stop_score;
  int row0_mode;           // as MergeParams.row0_mode: which row is "index 0" of the class's NMS input
  float* pad_score;        // [P] padded == 2: score of index 0 (what the padded selection slots gather, :332-335)
  float4* pad_box;         // [P] its box (clipped)
};

struct NmsShared {
  SelectScratch<RPP_NMS_NT> sel;
  u64 chunk[RPP_NMS_CHUNK];
  u64 lkeys[RPP_LIST_SMEM];
  float4 cbox[RPP_NMS_NT];   // canonical boxes of the current group
  float carea[RPP_NMS_NT];
  float4 corig[RPP_NMS_NT];  // boxes as emitted
  int nkept;
  int nk_slot[2];            // kept count handed from tile t to tile t+1 (double-buffered: see hard_nms_consume)
  int done;
  int need_all;              // padded == 2: the class's padded slots can reach the output -> count past the threshold
  // followed in dynamic shared memory by: float4 kbox[M_lim] (kept, canonical), float karea[M_lim]
};
__device__ __forceinline__ float4* nms_kbox(NmsShared* sh) { return reinterpret_cast<float4*>(sh + 1); }
__device__ __forceinline__ float* nms_karea(NmsShared* sh, int M_lim) {
  return reinterpret_cast<float*>(nms_kbox(sh) + M_lim);
}
__host__ __device__ static inline size_t nms_shared_bytes(int M_lim) { return sizeof(NmsShared) + (size_t)M_lim * 20 + 16; }

__device__ __forceinline__ float col_score(const ColProblemParams& P, float raw) {
  return P.is_logit ? sigmoid_f32(raw) : raw;
}

// MAPPED = false: the caller knows the rows are rows of the delta tensor itself (the per-class modes' warp probe)
template <bool MAPPED = true>
__device__ __forceinline__ float4 col_box(const ColProblemParams& P, int b, int c, u32 row) {
  if (P.boxes) {
    const int qi = P.q > 1 ? (c < P.q - 1 ? c : P.q - 1) : 0;  // boxes[:, min(q-1, c)] (:440)
    return P.boxes[((size_t)b * P.N + row) * P.q + qi];
  }
  // (one decode_box per call site: the row / delta source is selected first — three inlined copies of the binary64
  // exp code per site cost the soft per-class kernel 70 % of its speed)
  bool use_dlv = false;
  if (MAPPED) {
    if (P.row_keys) {
      row = key_tie(P.row_keys[(size_t)b * P.k_rows + row]) / (u32)P.C_src;
      use_dlv = true;
    } else if (P.dlv.L > 0) {
      use_dlv = true;
    }
  }
  float4 d;
  if (MAPPED && use_dlv) d = lv_delta(P.dlv, b, row); else d = lv_delta(P.lv, b, row);
  return decode_box(d, P.anchors[row], P.dp);
}

// Greedy hard NMS over one sorted chunk (m keys in sh->chunk).  The chunk is walked in groups of RPP_NMS_NT
// candidates (thread t owns candidate t: its box stays in registers) and each group in tiles of 32 = one warp:
//   (a) every unresolved candidate tests itself against the boxes kept since its last test (all warps busy);
//   (b) the tile's warp builds the 32x32 suppression mask among its still-alive candidates and resolves the
//       greedy order with a register bit-chain (no IoU on the serial path);
//   (c) the newly kept boxes are appended to the kept list; later tiles see them in their next (a).
// Equivalent to NonMaxSuppressionV5's hard branch / CombinedNMS's per-class loop: a candidate is kept iff no
// earlier kept box overlaps it by more than the threshold (the reverse-order early break of A.2 does not change
// the outcome).  Sets sh->done when M_lim are kept or k_lim candidates were consumed.
__device__ __forceinline__ bool iou_gt(float4 a, float area_a, float4 b, float area_b, float thr) {
  // degenerate boxes are stored as the empty box (inf, inf, -inf, -inf) with area 0: inter == 0 below (A.1)
  const float h0 = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float h1 = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = __fmul_rn(h0, h1);
  float iou = 0.0f;
  if (inter > 0.0f) iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return iou > thr;
}

// _bbox_overlap of tf.image.non_max_suppression_padded (image_ops_impl.py): no canonicalisation,
// inter / (area_a + area_b - inter + 1e-8) in fp32, and a box is suppressed when iou >= threshold.
__device__ __forceinline__ bool iou_padded_ge(float4 a, float area_a, float4 b, float area_b, float thr) {
  const float h0 = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float h1 = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = __fmul_rn(h1, h0);
  const float uni = __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-8f);
  return __fdiv_rn(inter, uni) >= thr;
}
template <bool PADDED>
__device__ __forceinline__ bool nms_suppresses(float4 a, float area_a, float4 b, float area_b, float thr) {
  return PADDED ? iou_padded_ge(a, area_a, b, area_b, thr) : iou_gt(a, area_a, b, area_b, thr);
}

// PADDED = true: the greedy scan that non_max_suppression_padded's tiled fixed-point iteration computes — same
// order (score desc, index asc), iou_padded_ge as the test, and a box whose coordinates are all <= 0 is never
// selected (TF counts `any(box > 0)`; such a box has IoU 0 with everything, so it does not suppress either).
template <bool PADDED>
__device__ void hard_nms_consume(const ColProblemParams& P, NmsShared* sh, int b, int c, size_t p, int m,
                                 long& consumed) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* kbox = nms_kbox(sh);
  float* karea = nms_karea(sh, P.M_lim);
  const long room = P.k_lim - consumed;
  int m_eff = (long)m < room ? m : (int)room;
  const float thr = P.iou_threshold;
  bool cut = false;
  if (P.pass == 2) {  // candidates below the image's bound can never reach the final top-M: stop there
    const float L = P.stop_L[b];
    int ok = 0;
    for (int i0 = 0; i0 < m_eff; i0 += RPP_NMS_NT)
      ok += __syncthreads_count(i0 + tid < m_eff && key_score(sh->chunk[i0 + tid]) >= L);
    cut = ok < m_eff;
    m_eff = ok;
  }
  if (PADDED && P.padded == 2) {
    if (consumed == 0 && P.row0_mode == 1) {
      // the per-class top-k ran first: index 0 of this class's NMS input is the head of the sorted stream
      if (tid == 0) {
        const u64 k0 = sh->chunk[0];
        float4 b0 = col_box(P, b, c, key_tie(k0));
        if (P.clip_before) b0 = clip01(b0);
        P.pad_score[p] = key_score(k0);
        P.pad_box[p] = b0;
        sh->need_all = key_score(k0) > P.stop_score;
      }
      __syncthreads();
    }
    if (!sh->need_all) {   // nothing at or below the score threshold can reach the output: stop there
      int ok = 0;
      for (int i0 = 0; i0 < m_eff; i0 += RPP_NMS_NT)
        ok += __syncthreads_count(i0 + tid < m_eff && key_score(sh->chunk[i0 + tid]) > P.stop_score);
      cut = cut || ok < m_eff;
      m_eff = ok;
    }
  }
  for (int g0 = 0; g0 < m_eff; g0 += RPP_NMS_NT) {
    const int gcount = m_eff - g0 < RPP_NMS_NT ? m_eff - g0 : RPP_NMS_NT;
    bool alive = tid < gcount;
    float4 bx = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    float area = 0.0f;
    if (alive) {
      float4 orig = col_box(P, b, c, key_tie(sh->chunk[g0 + tid]));
      if (P.clip_before) orig = clip01(orig);
      sh->corig[tid] = orig;
      if (PADDED) {
        bx = orig;
        area = __fmul_rn(__fsub_rn(orig.z, orig.x), __fsub_rn(orig.w, orig.y));
        alive = orig.x > 0.0f || orig.y > 0.0f || orig.z > 0.0f || orig.w > 0.0f;
      } else {
        const float4 cb = canon_box(orig, area);
        if (area > 0.0f) bx = cb; else area = 0.0f;
      }
      sh->cbox[tid] = bx;
      sh->carea[tid] = area;
    }
    // Kept count at the start of this group.  Read BEFORE the barrier: inside the tile loop the count travels
    // through nk_slot[], written by the warp of tile t before the loop's barrier and read by everybody after it,
    // so no thread ever reads a count in the same barrier interval in which another warp writes it.
    int nk = sh->nkept;
    __syncthreads();
    // every warp builds the suppression mask of its own tile now (pairwise IoU does not depend on what is kept):
    // bit j of `row` = candidate j < lane of my tile overlaps me.  All warps are busy; the serial part of a round
    // is then only the bit-chain.
    u32 row = 0u;
    {
      const int tbase = warp * 32;
      const int tcount = gcount - tbase < 32 ? gcount - tbase : 32;
      for (int j = 0; j < tcount - 1; ++j) {
        const float4 ob = sh->cbox[tbase + j];
        const float oa = sh->carea[tbase + j];
        if (j < lane && lane < tcount && nms_suppresses<PADDED>(bx, area, ob, oa, thr)) row |= 1u << j;
      }
    }
    int tested = 0;
    bool full = false;
    const int ntiles = (gcount + 31) >> 5;
    for (int tile = 0; tile < ntiles; ++tile) {
      if (alive && warp >= tile) {
        for (int k = tested; k < nk; ++k)
          if (nms_suppresses<PADDED>(bx, area, kbox[k], karea[k], thr)) { alive = false; break; }
      }
      tested = nk;
      if (warp == tile) {
        const u32 cand_bits = __ballot_sync(RPP_FULL_MASK, alive);
        u32 kept_bits = 0u;
#pragma unroll
        for (int l = 0; l < 32; ++l) {
          const u32 r = __shfl_sync(RPP_FULL_MASK, row, l);
          if (((cand_bits >> l) & 1u) && (r & kept_bits) == 0u) kept_bits |= 1u << l;
        }
        int nnew = __popc(kept_bits);
        const int room_k = P.M_cap - nk;
        while (nnew > room_k) {  // keep only the first room_k
          kept_bits &= ~(1u << (31 - __clz(kept_bits)));
          --nnew;
        }
        if ((kept_bits >> lane) & 1u) {
          const int pos = nk + __popc(kept_bits & ((1u << lane) - 1u));
          kbox[pos] = bx;
          karea[pos] = area;
          P.sel_key[p * P.M + pos] = sh->chunk[g0 + tid];
          P.sel_box[p * P.M + pos] = sh->corig[tid];
        }
        if (lane == 0) {
          sh->nk_slot[(tile + 1) & 1] = nk + nnew;
          sh->nkept = nk + nnew;
          if (nk + nnew >= P.M_cap) sh->done = 1;
        }
      }
      __syncthreads();
      nk = sh->nk_slot[(tile + 1) & 1];
      if (nk >= P.M_cap) { full = true; break; }
    }
    if (full) break;
  }
  consumed += m_eff;
  if (consumed >= P.k_lim || cut) {
    __syncthreads();
    if (tid == 0) sh->done = 1;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Argmax-iterate hard NMS over one UNSORTED chunk (finish pass of the hard per-class modes).  The classes the finish
// pass re-runs are the hot ones of a trained detector: hundreds to thousands of anchors of a few objects above the
// bound, nearly all of them suppressed by the first box of their object, a handful kept.  Sorting them and resolving
// them tile by tile (hard_nms_consume) spends its time on candidates that die anyway; here every round
//   * the block takes the best alive key (the next box greedy NMS keeps — same total order, same result),
//   * every alive candidate tests itself against that ONE box and tracks the best survivor of its thread,
// so a round is one pass over the chunk's alive slots and two barriers, and the number of rounds is the number of
// boxes kept.  keys[0..m): scored keys (0 = dead); abox[0..m): their boxes as emitted (decoded here).  Stops (sh->done)
// at M_cap kept boxes or at the first best key below the image's bound L; returns with the chunk exhausted otherwise.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 canon_or_empty(float4 orig, float& area) {
  const float4 cb = canon_box(orig, area);
  if (area > 0.0f) return cb;
  area = 0.0f;
  return make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
}

// prepared: abox[] already holds the boxes and every candidate was already tested against the kept boxes.
__device__ void hard_nms_argmax(const ColProblemParams& P, NmsShared* sh, float4* abox, u64* keys, int m, int b, int c,
                                size_t p, float L, bool prepared) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* kbox = nms_kbox(sh);
  float* karea = nms_karea(sh, P.M_lim);
  const float thr = P.iou_threshold;
  int nk = sh->nkept;   // (uniform: written before the barrier that precedes this call)
  // boxes of the chunk; candidates that a box kept earlier (previous chunks) suppresses die right away
  u64 best = 0ull;
  int best_i = -1;
  for (int i = tid; i < m; i += RPP_NMS_NT) {
    const u64 k = keys[i];
    if (k == 0ull) continue;
    if (prepared) {
      if (k > best) { best = k; best_i = i; }
      continue;
    }
    float4 orig = col_box(P, b, c, key_tie(k));
    if (P.clip_before) orig = clip01(orig);
    abox[i] = orig;
    float area;
    const float4 bx = canon_or_empty(orig, area);
    bool alive = true;
    for (int q = 0; q < nk; ++q)
      if (iou_gt(bx, area, kbox[q], karea[q], thr)) { alive = false; break; }
    if (!alive) { keys[i] = 0ull; continue; }
    if (k > best) { best = k; best_i = i; }
  }
  u64* red = reinterpret_cast<u64*>(sh->carea);   // [RPP_NMS_NT / 32] per-warp maxima (carea is free in this consumer)
  for (;;) {
    u64 wb = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u64 t = __shfl_xor_sync(RPP_FULL_MASK, wb, o);
      wb = t > wb ? t : wb;
    }
    __syncthreads();   // the previous round's readers of red[] / kbox[nk - 1] are through
    if (lane == 0) red[warp] = wb;
    __syncthreads();
    u64 bk = red[0];
#pragma unroll
    for (int w = 1; w < RPP_NMS_NT / 32; ++w) bk = red[w] > bk ? red[w] : bk;
    if (bk == 0ull) return;                                  // chunk exhausted
    if (key_score(bk) < L) {                                 // nothing below the image's bound can reach the top M
      if (tid == 0) sh->done = 1;
      __syncthreads();
      return;
    }
    if (best == bk) {                                        // exactly one owner: keys are unique
      const float4 orig = abox[best_i];
      float area;
      kbox[nk] = canon_or_empty(orig, area);
      karea[nk] = area;
      P.sel_key[p * P.M + nk] = bk;
      P.sel_box[p * P.M + nk] = orig;
      keys[best_i] = 0ull;
      sh->nkept = nk + 1;
      if (nk + 1 >= P.M_cap) sh->done = 1;
    }
    __syncthreads();
    ++nk;
    if (nk >= P.M_cap) return;
    const float4 nb = kbox[nk - 1];
    const float na = karea[nk - 1];
    best = 0ull;
    best_i = -1;
    for (int i = tid; i < m; i += RPP_NMS_NT) {
      const u64 k = keys[i];
      if (k == 0ull) continue;
      float area;
      const float4 bx = canon_or_empty(abox[i], area);
      if (iou_gt(bx, area, nb, na, thr)) { keys[i] = 0ull; continue; }
      if (k > best) { best = k; best_i = i; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Soft NMS consumer: NonMaxSuppressionV5 with soft_nms_sigma > 0, lazily re-scored exactly as the
// TF kernel does it.  The priority queue is split in two: candidates never popped yet are the not-yet-consumed part
// of the sorted stream (their order is static), and candidates popped, decayed and pushed back live in R (shared
// memory, spilling to global).  Each step pops the larger of (stream head, max of R); a popped candidate multiplies
// its score by the weights of the boxes selected since its last visit, newest first, in fp32 in exactly that order,
// stopping when it falls to the score threshold; it is selected iff the score did not change.
// expf_glibc reproduces libm's expf bit for bit (checked against glibc on 4.5e8 inputs): TF's kernel calls
// Eigen::numext::exp<float> = expf.
// ---------------------------------------------------------------------------------------------------------------
// (the 32-entry 2^(i/32) table of glibc's expf — sysdeps/ieee754/flt-32/e_exp2f_data.c, LGPL-2.1+ — restated: these
// are the correctly rounded binary64 values of 2^(i/32) with the exponent bits adjusted, i.e. mathematical constants)
__constant__ u64 c_exp2f_tab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

// `tab` = c_exp2f_tab or a copy of it: the index differs per lane, and the constant cache serves one address per
// cycle, so warps that evaluate many weights at once read a shared-memory copy instead.
__device__ __forceinline__ float expf_glibc_tab(float x, const u64* tab) {
  if (!(x > -87.0f && x < 88.0f)) return (float)exp((double)x);  // under/overflow tails: correctly rounded exp
  const double N = 32.0;
  const double InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
  const double z = __dmul_rn(InvLn2N, (double)x);
  double kd = __dadd_rn(z, SHIFT);
  const u64 ki = (u64)__double_as_longlong(kd);
  kd = __dsub_rn(kd, SHIFT);
  const double r = __dsub_rn(z, kd);
  const u64 t = tab[ki & 31u] + (ki << 47);
  const double sc = __longlong_as_double((long long)t);
  const double zz = __dadd_rn(__dmul_rn(C0, r), C1);
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
  y = __dadd_rn(__dmul_rn(zz, r2), y);
  y = __dmul_rn(y, sc);
  return __double2float_rn(y);
}
__device__ __forceinline__ float expf_glibc(float x) { return expf_glibc_tab(x, c_exp2f_tab); }

#define RPP_SOFT_RS 512

struct SoftShared {
  u64 rkey[RPP_SOFT_RS];
  uint2 rmeta[RPP_SOFT_RS];   // {suppress_begin_index, row}
  float4 rbox[RPP_SOFT_RS];   // canonical box (the empty box when degenerate)
  int rcount;
};

struct RStore {
  SoftShared* s;
  u64* gk; uint2* gm; float4* gb;
  __device__ __forceinline__ u64 key(int i) const { return i < RPP_SOFT_RS ? s->rkey[i] : gk[i - RPP_SOFT_RS]; }
  __device__ __forceinline__ uint2 meta(int i) const { return i < RPP_SOFT_RS ? s->rmeta[i] : gm[i - RPP_SOFT_RS]; }
  __device__ __forceinline__ float4 box(int i) const { return i < RPP_SOFT_RS ? s->rbox[i] : gb[i - RPP_SOFT_RS]; }
  __device__ __forceinline__ void set(int i, u64 k, uint2 m, float4 b) {
    if (i < RPP_SOFT_RS) { s->rkey[i] = k; s->rmeta[i] = m; s->rbox[i] = b; }
    else { gk[i - RPP_SOFT_RS] = k; gm[i - RPP_SOFT_RS] = m; gb[i - RPP_SOFT_RS] = b; }
  }
};

__device__ __forceinline__ float iou_val(float4 a, float area_a, float4 b, float area_b) {
  const float h0 = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float h1 = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = __fmul_rn(h0, h1);
  if (!(inter > 0.0f)) return 0.0f;
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// m > 0: consume a sorted chunk of the stream (returns when it is exhausted or the problem is done);
// m == 0 && final: the stream is over, drain R.
__device__ void soft_nms_consume(const ColProblemParams& P, NmsShared* sh, SoftShared* ss, int b, int c, size_t p,
                                 int m, long& consumed, bool final) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* kbox = nms_kbox(sh);
  float* karea = nms_karea(sh, P.M_lim);
  RStore R{ss, P.r_key + p * (size_t)P.r_cap, P.r_meta + p * (size_t)P.r_cap, P.r_box + p * (size_t)P.r_cap};
  const long room = P.k_lim - consumed;
  const int m_eff = (long)m < room ? m : (int)room;
  const float thr = P.score_threshold;
  const int ngroups = final ? 1 : (m_eff + RPP_NMS_NT - 1) / RPP_NMS_NT;
  for (int g = 0; g < ngroups; ++g) {
    const int g0 = g * RPP_NMS_NT;
    const int gcount = final ? 0 : (m_eff - g0 < RPP_NMS_NT ? m_eff - g0 : RPP_NMS_NT);
    if (tid < gcount) {
      float4 orig = col_box(P, b, c, key_tie(sh->chunk[g0 + tid]));
      if (P.clip_before) orig = clip01(orig);
      float area;
      const float4 cb = canon_box(orig, area);
      sh->corig[tid] = orig;
      sh->cbox[tid] = area > 0.0f ? cb : make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
      sh->carea[tid] = area > 0.0f ? area : 0.0f;
    }
    __syncthreads();
    if (warp == 0) {
      int nsel = sh->nkept;
      int rcount = ss->rcount;
      int pos = 0;
      for (;;) {
        if (nsel >= P.M_cap) { if (lane == 0) sh->done = 1; break; }
        // stream head
        u64 head = 0ull, head_cmp = 0ull;
        if (pos < gcount) {
          head = sh->chunk[g0 + pos];
          head_cmp = P.tie_is_rank ? make_key(key_score(head), (u32)(consumed + g0 + pos)) : head;
        }
        // max of R
        u64 best = 0ull;
        int best_i = -1;
        for (int i = lane; i < rcount; i += 32) {
          const u64 k = R.key(i);
          if (k > best) { best = k; best_i = i; }
        }
        u64 wbest = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const u64 other = __shfl_xor_sync(RPP_FULL_MASK, wbest, o);
          wbest = other > wbest ? other : wbest;
        }
        const u32 owner = __ballot_sync(RPP_FULL_MASK, best == wbest && best != 0ull);
        const int r_i = owner ? __shfl_sync(RPP_FULL_MASK, best_i, __ffs(owner) - 1) : -1;
        if (head == 0ull && (!final || wbest == 0ull)) break;  // need more stream / everything drained
        const bool from_stream = head != 0ull && head_cmp > wbest;
        if (P.pass == 2 && key_score(from_stream ? head_cmp : wbest) < P.stop_L[b]) {
          // the queue maximum is below the image's bound: nothing this class selects from now on can matter
          if (lane == 0) sh->done = 1;
          break;
        }
        float score, area;
        float4 box;
        u32 row, tie;
        int begin;
        if (from_stream) {
          score = key_score(head); row = key_tie(head); tie = key_tie(head_cmp); begin = 0;
          box = sh->cbox[pos]; area = sh->carea[pos];
        } else {
          score = key_score(wbest); tie = key_tie(wbest);
          const uint2 mt = R.meta(r_i);
          begin = (int)mt.x; row = mt.y;
          box = R.box(r_i);
          area = box.z > box.x ? __fmul_rn(__fsub_rn(box.z, box.x), __fsub_rn(box.w, box.y)) : 0.0f;
        }
        const float original = score;
        bool dropped = false;
        for (int j = nsel - 1; j >= begin && !dropped; j -= 32) {
          const int jj = j - lane;
          float w = 1.0f;
          if (jj >= begin) {
            const float sim = iou_val(box, area, kbox[jj], karea[jj]);
            // sim == 0 (no overlap, the common case): expf(scale * 0 * 0) = expf(0) = 1 exactly
            if (sim != 0.0f) w = expf_glibc(__fmul_rn(__fmul_rn(P.soft_scale, sim), sim));
            if (!P.soft_ignores_iou && sim > P.iou_threshold) w = 0.0f;
          }
          // multiply in the kernel's order (newest selected first = ascending lane); a weight of exactly 1.0 leaves
          // the score and the threshold test unchanged, so only the lanes that overlap are walked
          u32 nz = __ballot_sync(RPP_FULL_MASK, w != 1.0f);
          while (nz) {
            const int t = __ffs(nz) - 1;
            nz &= nz - 1u;
            score = __fmul_rn(score, __shfl_sync(RPP_FULL_MASK, w, t));
            if (score <= thr) { dropped = true; break; }
          }
        }
        if (from_stream) ++pos;
        __syncwarp();             // every lane has read its R entry before lane 0 rewrites R below
        if (score == original) {  // select
          if (lane == 0) {
            kbox[nsel] = box;
            karea[nsel] = area;
            P.sel_key[p * P.M + nsel] = make_key(score, row);
            float4 ob;
            if (from_stream) ob = sh->corig[pos - 1];
            else { ob = col_box(P, b, c, row); if (P.clip_before) ob = clip01(ob); }
            P.sel_box[p * P.M + nsel] = ob;
          }
          ++nsel;
          if (!from_stream) {  // remove from R (swap with last)
            --rcount;
            if (lane == 0 && r_i != rcount) R.set(r_i, R.key(rcount), R.meta(rcount), R.box(rcount));
          }
        } else if (!dropped && score > thr) {  // push back, re-scored
          const int slot = from_stream ? rcount : r_i;
          if (lane == 0) R.set(slot, make_key(score, tie), make_uint2((u32)nsel, row), box);
          if (from_stream) ++rcount;
        } else if (!from_stream) {  // fell to the threshold: gone
          --rcount;
          if (lane == 0 && r_i != rcount) R.set(r_i, R.key(rcount), R.meta(rcount), R.box(rcount));
        }
        __syncwarp();
      }
      __syncwarp();   // every lane has read the counts of this group before lane 0 replaces them
      if (lane == 0) { sh->nkept = nsel; ss->rcount = rcount; }
    }
    __syncthreads();
    if (sh->done) break;
  }
  if (!final) consumed += m_eff;  // at k_lim the caller stops the stream and drains R
}

// Top-k emission consumer (FilterTopKDetections): the stream IS the sorted top-k.
__device__ void emit_consume(const ColProblemParams& P, NmsShared* sh, size_t p, int m, long& consumed) {
  const long room = P.k_lim - consumed;
  const int m_eff = (long)m < room ? m : (int)room;
  for (int i = threadIdx.x; i < m_eff; i += RPP_NMS_NT) P.emit_key[p * (size_t)P.k_lim + consumed + i] = sh->chunk[i];
  consumed += m_eff;
  if (consumed >= P.k_lim) {
    __syncthreads();
    if (threadIdx.x == 0) sh->done = 1;
    __syncthreads();
  }
}

#define RPP_CONSUME_HARD 0
#define RPP_CONSUME_SOFT 1
#define RPP_CONSUME_EMIT 2
#define RPP_CONSUME_PADDED 3   // hard NMS with tf.image.non_max_suppression_padded semantics (TPU branches)

template <int MODE>
__device__ __forceinline__ void col_problem_body(const ColProblemParams& P, const size_t p, NmsShared* sh,
                                                 SoftShared* ss) {
  const int tid = threadIdx.x;
  const int b = (int)(p / P.C), c = (int)(p % P.C);
  if (MODE == RPP_CONSUME_EMIT && P.emit_done && P.emit_done[p]) return;   // done by emit_sort_kernel
  if (MODE != RPP_CONSUME_EMIT && P.pass == 2) {
    const float bd = P.bound[p];
    if (bd == -INFINITY || bd < P.stop_L[b]) return;   // the probe already holds everything that can matter
  }
  if (tid == 0) {
    sh->nkept = 0;
    sh->done = 0;
    sh->need_all = 0;
    sh->nk_slot[0] = 0;
    sh->nk_slot[1] = 0;
    if (MODE == RPP_CONSUME_SOFT) ss->rcount = 0;
    if (MODE == RPP_CONSUME_PADDED && P.padded == 2) {
      float s0 = -INFINITY;
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (P.row0_mode == 0) {   // index 0 of the NMS input = row 0 of the source
        s0 = col_score(P, lv_val(P.lv, b, 0, P.C, c));
        b0 = col_box(P, b, c, 0u);
        if (P.clip_before) b0 = clip01(b0);
        sh->need_all = s0 > P.stop_score;
      }
      P.pad_score[p] = s0;
      P.pad_box[p] = b0;
    }
  }
  __syncthreads();

  long consumed = 0;
  auto consume = [&](int m) {
    if (MODE == RPP_CONSUME_HARD) hard_nms_consume<false>(P, sh, b, c, p, m, consumed);
    else if (MODE == RPP_CONSUME_PADDED) hard_nms_consume<true>(P, sh, b, c, p, m, consumed);
    else if (MODE == RPP_CONSUME_SOFT) soft_nms_consume(P, sh, ss, b, c, p, m, consumed, false);
    else emit_consume(P, sh, p, m, consumed);
  };
  const int want0 = MODE == RPP_CONSUME_EMIT ? RPP_NMS_CHUNK : P.want0;

  const float T = P.T[p];
  u32 n_raw = P.cand_count[p];
  const bool converted = (n_raw & 0x80000000u) != 0u;   // a previous pass left u64 keys in the list
  n_raw &= 0x7fffffffu;
  const bool overflow = n_raw > (u32)P.CAP;
  int n_list = (overflow || P.force_scan) ? 0 : (int)n_raw;
  const bool list_complete = !(T > P.T_min);  // the list holds every element above the score threshold
  float s_edge = P.score_threshold;
  if (!list_complete) s_edge = col_score(P, T);
  if (overflow || P.force_scan) s_edge = INFINITY;

  // ---- phase A0: the head of the list, selected on RAW logits ------------------------------------------------
  // The consumer usually wants a few dozen candidates, so evaluating the binary64 sigmoid for the whole list is
  // wasted work.  The score is monotone in the logit: the top-`want` of the list by (logit desc, index asc) is a
  // complete prefix of the score order for every score strictly above S(lowest selected logit) =: e0 (the same
  // "edge rule" as for the collect threshold, one level down).  Only those are scored, re-keyed by (score, index),
  // sorted and consumed here; if the consumer wants more, phase A continues below the bound e0 with the full list.
  u64 KB_A = ~0ull;          // phase A consumes keys below this bound
  bool skip_A = false;
  // ---- finish pass of the hard modes: argmax-iterate (hard_nms_argmax) over what the probe's boxes leave alive -----
  // The boxes the probe kept ARE the first boxes greedy NMS keeps, so the finish pass continues from them instead of
  // starting over: ONE pass over the list scores every candidate at or above the bound (a logit clearly below the
  // bound's pre-image is dropped without evaluating the sigmoid), decodes it, tests it against the probe's boxes and
  // stages the survivors — on trained-detector inputs a few per cent of a hot class's list — in shared memory, where
  // the argmax rounds run.  More survivors than shared memory holds (a class whose candidates mostly do NOT overlap:
  // the case the tile NMS below is built for) leaves the class to the generic path, from scratch.
  bool argmax_done = false;
  if (MODE == RPP_CONSUME_HARD && P.pass == 2 && P.argmax && n_list > 0 && (long)n_list <= P.k_lim) {
    float4* abox = reinterpret_cast<float4*>(ss);
    float4* kbox = nms_kbox(sh);
    float* karea = nms_karea(sh, P.M_lim);
    u64* pk = sh->chunk;   // keys of the probe's boxes
    const float L = P.stop_L[b];
    const float thr = P.iou_threshold;
    const uint2* lst = P.cand + p * (size_t)P.CAP;
    const u64* gkeys = reinterpret_cast<const u64*>(lst);
    int nk0 = P.sel_cnt[p];
    if (nk0 > P.M_cap) nk0 = P.M_cap;
    if (tid < nk0) {
      float area;
      kbox[tid] = canon_or_empty(P.sel_box[p * P.M + tid], area);
      karea[tid] = area;
      pk[tid] = P.sel_key[p * P.M + tid];
    }
    float raw_lo = -INFINITY;
    if (P.is_logit && L > 0.0f && L < 1.0f) {
      const float x = __logf(L / (1.0f - L));
      raw_lo = x - 1e-3f * (1.0f + fabsf(x));
    }
    // The probe's boxes must all come out of the LIST: a probe that went on into the exact scan below the list's edge
    // kept boxes the scan of this pass would meet again (a zero-area box, or an IoU threshold of 1, is not suppressed
    // by its own copy) — such a class starts over on the generic path.
    const bool from_list =
        __syncthreads_and(tid >= nk0 || list_complete || key_score(P.sel_key[p * P.M + tid]) > s_edge) != 0;
    // `below` = some consumable candidate was dropped for being below L (the class is then finished once the list is)
    int below = 0, n_valid = 0;
    // (four candidates per thread and trip: the list entries, then the deltas / anchors of the live ones, are loaded
    // together — one candidate at a time the loop was a chain of three dependent global loads per candidate, and this
    // kernel has 16 warps per SM to hide them with)
    const bool plain_boxes = !P.boxes && !P.row_keys && P.dlv.L == 0;   // rows of the delta tensor itself
    if (from_list)
    for (int i0 = tid; i0 < n_list; i0 += 4 * RPP_NMS_NT) {
      u64 k4[4];
      if (converted) {
#pragma unroll
        for (int u = 0; u < 4; ++u) k4[u] = i0 + u * RPP_NMS_NT < n_list ? gkeys[i0 + u * RPP_NMS_NT] : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k4[u] != 0ull && key_score(k4[u]) < L) { k4[u] = 0ull; below = 1; }
      } else {
        uint2 e4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          e4[u] = i0 + u * RPP_NMS_NT < n_list ? lst[i0 + u * RPP_NMS_NT] : make_uint2(0xff800000u /* -inf */, 0u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          k4[u] = 0ull;
          if (i0 + u * RPP_NMS_NT >= n_list) continue;
          const float raw = __uint_as_float(e4[u].x);
          if (P.is_logit && raw < raw_lo) {
            below = 1;   // (its score is below L; whether it was consumable at all does not matter: it only cuts)
          } else {
            const float s = col_score(P, raw);
            if (s > P.score_threshold && (list_complete || s > s_edge)) {
              if (s < L) below = 1; else k4[u] = make_key(s, e4[u].y);
            }
          }
        }
      }
      float4 d4[4], a4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k4[u] == 0ull) continue;
        ++n_valid;
        bool own = false;
        for (int q = 0; q < nk0; ++q) own = own || pk[q] == k4[u];   // one of the probe's own boxes
        if (own) { k4[u] = 0ull; continue; }
        if (plain_boxes) {
          const u32 row = key_tie(k4[u]);
          d4[u] = lv_delta(P.lv, b, row);
          a4[u] = P.anchors[row];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u64 k = k4[u];
        if (k == 0ull) continue;
        float4 orig = plain_boxes ? decode_box(d4[u], a4[u], P.dp) : col_box(P, b, c, key_tie(k));
        if (P.clip_before) orig = clip01(orig);
        float area;
        const float4 bx = canon_or_empty(orig, area);
        bool dead = false;
        for (int q = 0; q < nk0; ++q)
          if (iou_gt(bx, area, kbox[q], karea[q], thr)) { dead = true; break; }
        if (dead) continue;
        const int slot = atomicAdd(&sh->nk_slot[1], 1);
        if (slot < RPP_LIST_SMEM) { sh->lkeys[slot] = k; abox[slot] = orig; }
      }
    }
    if (n_valid) atomicAdd(&sh->nk_slot[0], n_valid);   // (both counters are zeroed with the per-problem state above)
    below = __syncthreads_or(below);
    n_valid = sh->nk_slot[0];
    const int nsurv = sh->nk_slot[1];
    if (from_list && nsurv <= RPP_LIST_SMEM) {
      if (tid == 0) { sh->nkept = nk0; if (nk0 >= P.M_cap) sh->done = 1; }
      __syncthreads();
      if (nk0 < P.M_cap) hard_nms_argmax(P, sh, abox, sh->lkeys, nsurv, b, c, p, L, /*prepared=*/true);
      consumed += n_valid;
      // the list is exhausted: what was left out of it scores at most s_edge — if that (or anything dropped above)
      // is below the bound, the class is finished; otherwise phase B continues below the edge
      if (!sh->done && (below || list_complete || s_edge < L)) {
        __syncthreads();
        if (tid == 0) sh->done = 1;
      }
      __syncthreads();
      skip_A = true;
      argmax_done = true;
    }
    __syncthreads();
  }
  if (argmax_done) {
  } else
  if (MODE != RPP_CONSUME_EMIT && P.is_logit && n_list > 0 && n_list <= RPP_LIST_SMEM && !converted) {
    const uint2* lst = P.cand + p * (size_t)P.CAP;
    for (int i = tid; i < n_list; i += RPP_NMS_NT) {
      const uint2 e = lst[i];
      sh->lkeys[i] = ((u64)ord_f32(__uint_as_float(e.x)) << 32) | (u64)(0xffffffffu - e.y);
    }
    __syncthreads();
    u64 KBr = ~0ull;
    const int m = select_chunk<RPP_NMS_NT>([&](int i) { return sh->lkeys[i]; }, n_list, KBr, want0, sh->chunk,
                                           RPP_NMS_CHUNK, &sh->sel, /*sort=*/false);
    // m >= 1 (the list is not empty).  Everything outside the chunk has a raw key < KBr (the cut), i.e. a logit <=
    // the float encoded in the cut's upper half.
    const bool whole = m == n_list;
    float e0;
    if (whole) e0 = list_complete ? -INFINITY : s_edge;
    else e0 = sigmoid_f32(unord_f32((u32)(KBr >> 32)));
    const int P2 = next_pow2(m < 2 ? 2 : m);
    for (int i = tid; i < P2; i += RPP_NMS_NT) {
      u64 k = 0ull;
      if (i < m) {
        const u64 rk = sh->chunk[i];
        const float sc = sigmoid_f32(unord_f32((u32)(rk >> 32)));
        if (sc > P.score_threshold && sc > e0) k = make_key(sc, key_tie(rk));
      }
      sh->chunk[i] = k;
    }
    __syncthreads();
    bitonic_sort_desc<RPP_NMS_NT>(sh->chunk, P2);   // true order: (score desc, index asc); invalid keys sink
    int mv = 0;
    for (int i0 = 0; i0 < m; i0 += RPP_NMS_NT) mv += __syncthreads_count(i0 + tid < m && sh->chunk[i0 + tid] != 0ull);
    if (mv > 0) consume(mv);
    if (whole) skip_A = true;                       // nothing of the list is left that phase A may consume
    else KB_A = (u64)(ord_f32(e0) + 1u) << 32;      // phase A: scores <= e0
  }
  // ---- phase A: the collected list, keyed by score ----------------------------------------------------------
  if (n_list > 0 && !skip_A && !sh->done && consumed < P.k_lim) {
    uint2* lst = P.cand + p * (size_t)P.CAP;
    u64* gkeys = reinterpret_cast<u64*>(lst);
    u64* keys = n_list <= RPP_LIST_SMEM ? sh->lkeys : gkeys;
    if (!converted) {
      for (int i = tid; i < n_list; i += RPP_NMS_NT) {
        const uint2 e = lst[i];
        const float s = col_score(P, __uint_as_float(e.x));
        // consumable now: strictly above everything that was NOT collected (those score <= s_edge)
        const bool ok = s > P.score_threshold && (list_complete || s > s_edge);
        keys[i] = ok ? make_key(s, e.y) : 0ull;
      }
    }
    __syncthreads();
    // long lists are converted in place (global memory): remember it for the finish pass.  (After the barrier: every
    // thread has derived `converted` from the count word by now.)
    if (!converted && keys == gkeys && tid == 0) P.cand_count[p] = n_raw | 0x80000000u;
    u64 KB = KB_A;
    int want = want0;
    while (!sh->done && consumed < P.k_lim) {
      const int m = select_chunk<RPP_NMS_NT>([&](int i) { return keys[i]; }, n_list, KB, want, sh->chunk,
                                             RPP_NMS_CHUNK, &sh->sel);
      if (m == 0) break;
      consume(m);
      want = RPP_NMS_CHUNK;
    }
  }
  // ---- phase B: exact scan of the column for everything at or below the edge ---------------------------------
  if (!sh->done && consumed < P.k_lim && (!list_complete || overflow || P.force_scan)) {
    if (tid == 0 && P.scan_count) atomicAdd(P.scan_count, 1u);
    u64 KB = (s_edge == INFINITY) ? ~0ull : ((u64)(ord_f32(s_edge) + 1u) << 32);
    auto keyfn = [&](int i) -> u64 {
      const float raw = lv_val(P.lv, b, i, P.C, c);
      if (!(raw >= P.T_min)) return 0ull;
      const float s = col_score(P, raw);
      return s > P.score_threshold ? make_key(s, (u32)i) : 0ull;
    };
    int want = want0;
    while (!sh->done && consumed < P.k_lim) {
      const int m = select_chunk<RPP_NMS_NT>(keyfn, (int)P.N, KB, want, sh->chunk, RPP_NMS_CHUNK, &sh->sel);
      if (m == 0) break;
      consume(m);
      want = RPP_NMS_CHUNK;
    }
  }
  if (MODE == RPP_CONSUME_SOFT) {
    if (!sh->done) soft_nms_consume(P, sh, ss, b, c, p, 0, consumed, true);  // stream over: drain the queue
  }
  if (MODE != RPP_CONSUME_EMIT && tid == 0) {
    const int nk = sh->nkept;
    P.sel_cnt[p] = nk;
    if (P.pass == 1)   // stopped by the probe cap: later boxes of this class score <= the last kept one
      P.bound[p] = (nk >= P.M_cap && nk > 0) ? key_score(P.sel_key[p * P.M + nk - 1]) : -INFINITY;
  }
}

template <int MODE>
__global__ void __launch_bounds__(RPP_NMS_NT) col_problem_kernel(ColProblemParams P) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsShared* sh = reinterpret_cast<NmsShared*>(smem_raw);
  SoftShared* ss = reinterpret_cast<SoftShared*>(smem_raw + ((nms_shared_bytes(P.M_lim) + 15) & ~(size_t)15));
  __shared__ u32 s_item;
  for (u32 it = 0;; ++it) {   // one problem per block, or a persistent block popping the finish pass's worklist
    size_t p = blockIdx.x;
    if (P.work_items) {
      __syncthreads();        // the previous problem is finished by every thread (and s_item was read)
      if (threadIdx.x == 0) s_item = atomicAdd(&P.work_ctl[1], 1u);
      __syncthreads();
      if (s_item >= P.work_ctl[0]) break;
      p = P.work_items[s_item];
    } else if (P.n_loop > 0) {
      p = blockIdx.x + (size_t)it * gridDim.x;
      if ((long)p >= P.n_loop) break;
      if (it > 0) __syncthreads();
    } else if (it > 0) {
      break;
    }
    col_problem_body<MODE>(P, p, sh, ss);
  }
}

// Per image: stop_L = the Mtop-th best score among the boxes the probes kept (-inf if there are fewer): every one of
// them is a real final candidate, so the image's Mtop-th best FINAL score is >= stop_L.
__global__ void perclass_bound_kernel(const u64* __restrict__ sel_key, const int* __restrict__ sel_cnt, int C, int M,
                                      int m1, int Mtop, float* __restrict__ stop_L, const float* __restrict__ bound,
                                      u32* __restrict__ work_items, u32* __restrict__ work_ctl) {
  pdl_enter();
  extern __shared__ float s_sc[];  // [C * m1]
  __shared__ int s_n;
  __shared__ float s_L;
  const int b = blockIdx.x;
  const int n_all = C * m1;
  for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
    const int c = i / m1, slot = i - c * m1;
    s_sc[i] = slot < sel_cnt[(size_t)b * C + c] ? key_score(sel_key[((size_t)b * C + c) * M + slot]) : -INFINITY;
  }
  if (threadIdx.x == 0) { s_n = 0; s_L = -INFINITY; }
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < n_all; i += blockDim.x) local += s_sc[i] > -INFINITY;
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  if (s_n >= Mtop) {
    // rank by counting, 4 lanes per element (the counting loop is this kernel's latency chain)
    for (int i0 = 0; i0 < n_all; i0 += blockDim.x / 4) {
      const int i = i0 + (threadIdx.x >> 2), sub = threadIdx.x & 3;
      const float v = i < n_all ? s_sc[i] : -INFINITY;
      int rank = 0;
      if (v > -INFINITY)
        for (int j = sub; j < n_all; j += 4) {
          const float o = s_sc[j];
          rank += (o > v) || (o == v && j < i);
        }
      rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 1);
      rank += __shfl_xor_sync(RPP_FULL_MASK, rank, 2);
      if (v > -INFINITY && sub == 0 && rank == Mtop - 1) s_L = v;   // exactly one element has this rank
    }
  }
  __syncthreads();
  const float L = s_L;
  if (threadIdx.x == 0) stop_L[b] = L;
  // worklist of the finish pass: the classes whose probe stopped at its cap with a bound that can still matter
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float bd = bound[(size_t)b * C + c];
    if (!(bd == -INFINITY || bd < L)) work_items[atomicAdd(&work_ctl[0], 1u)] = (u32)((size_t)b * C + c);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-per-problem PROBE of the hard per-class modes (pass 1 of the cross-class bound).  The probe only needs the
// first few boxes of a class, so it avoids the block machinery altogether: no block barrier, one warp = one problem.
//   1. every lane scans its share of the candidate list keeping its 3 largest raw keys (logit bits | ~index);
//   2. tau = the largest 3rd-best over the lanes: every key > tau is among some lane's best two, so {key > tau} is a
//      COMPLETE prefix of the list in raw order (<= 64 keys); everything else scores <= e0 = score(logit(tau));
//   3. the prefix is scored (sigmoid only here), re-keyed by (score, index), candidates not strictly above e0 dropped
//      (same edge rule as everywhere), and sorted with a 64-key register bitonic network (shuffles);
//   4. greedy NMS over up to two tiles of 32 with the suppression bit-mask / bit-chain, stopping at M_cap boxes.
// bound[p] = score of the last kept box when the cap was hit, else an upper bound for anything the class can still
// keep (e0, the collect edge, or -inf when the class is exhausted).  Exactness never depends on tau.
// ---------------------------------------------------------------------------------------------------------------
// 4 warps per block, at least 9 blocks per SM (<= 56 registers): all 5 120 problems of configs[1] are resident at once
// (8-warp blocks at 50 registers left 48 of the 640 blocks to a second wave)
#ifndef RPP_PROBE_WARPS
#define RPP_PROBE_WARPS 4
#endif
#ifndef RPP_PROBE_MINB
#define RPP_PROBE_MINB 9
#endif
#define RPP_PROBE_MAXCAP 16

struct ProbeWarpShared {
  float4 kbox[RPP_PROBE_MAXCAP];
  float karea[RPP_PROBE_MAXCAP];
};

__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const u64 t = __shfl_xor_sync(RPP_FULL_MASK, v, o);
    v = t > v ? t : v;
  }
  return v;
}

__global__ void __launch_bounds__(RPP_PROBE_WARPS * 32, RPP_PROBE_MINB) probe_warp_kernel(ColProblemParams P, size_t n_problems) {
  pdl_enter();
  __shared__ ProbeWarpShared s_all[RPP_PROBE_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t p = (size_t)blockIdx.x * RPP_PROBE_WARPS + warp;
  if (p >= n_problems) return;
  ProbeWarpShared* sh = &s_all[warp];
  const int b = (int)(p / P.C), c = (int)(p % P.C);
  const u32 n_raw = P.cand_count[p];
  if (P.force_scan || n_raw > (u32)P.CAP) {   // no usable list: the finish pass does the whole class
    if (lane == 0) { P.sel_cnt[p] = 0; P.bound[p] = INFINITY; }
    return;
  }
  const int n = (int)n_raw;
  const float T = P.T[p];
  const bool list_complete = !(T > P.T_min);
  const float s_edge = list_complete ? -INFINITY : col_score(P, T);
  const uint2* lst = P.cand + p * (size_t)P.CAP;

  // 1. per-lane top-3 raw keys
  u64 t0 = 0ull, t1 = 0ull, t2 = 0ull;
  for (int i0 = lane; i0 < n; i0 += 4 * 32) {   // 4 independent loads in flight per lane
    uint2 e4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) e4[u] = i0 + u * 32 < n ? lst[i0 + u * 32] : make_uint2(0u, 0u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const u64 rk = i0 + u * 32 < n
                         ? ((u64)ord_f32(__uint_as_float(e4[u].x)) << 32) | (u64)(0xffffffffu - e4[u].y)
                         : 0ull;
      if (rk > t2) {
        if (rk > t1) {
          t2 = t1;
          if (rk > t0) { t1 = t0; t0 = rk; } else { t1 = rk; }
        } else {
          t2 = rk;
        }
      }
    }
  }
  // 2. complete prefix {rk > tau}
  const u64 tau = warp_max_u64(t2);
  float e0 = s_edge;                                 // nothing outside the prefix scores above e0
  if (tau != 0ull) e0 = col_score(P, unord_f32((u32)(tau >> 32)));
  const bool whole_list = tau == 0ull;
  // 3. score + re-key the prefix (two slots per lane), drop what is not strictly above e0 / the score threshold
  u64 k[2];
  {
    const u64 r[2] = {t0, t1};
#pragma unroll
    for (int sidx = 0; sidx < 2; ++sidx) {
      k[sidx] = 0ull;
      if (r[sidx] > tau) {
        const float sc = col_score(P, unord_f32((u32)(r[sidx] >> 32)));
        if (sc > P.score_threshold && sc > e0) k[sidx] = make_key(sc, key_tie(r[sidx]));
      }
    }
  }
  // 64-key descending bitonic sort across the warp: element e = slot * 32 + lane
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        if (k[0] < k[1]) { const u64 t = k[0]; k[0] = k[1]; k[1] = t; }   // size == 64: all descending
      } else {
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
          const int e = sidx * 32 + lane;
          const u64 other = __shfl_xor_sync(RPP_FULL_MASK, k[sidx], j);
          const bool desc = (e & size) == 0;
          const bool low = (lane & j) == 0;
          const bool keep_max = desc == low;
          k[sidx] = keep_max ? (other > k[sidx] ? other : k[sidx]) : (other < k[sidx] ? other : k[sidx]);
        }
      }
    }
  }
  // pre_nms_top_k caps the candidates a class may consume
  if ((long)lane >= P.k_lim) k[0] = 0ull;
  if ((long)(32 + lane) >= P.k_lim) k[1] = 0ull;
  const int n_valid = __popc(__ballot_sync(RPP_FULL_MASK, k[0] != 0ull)) + __popc(__ballot_sync(RPP_FULL_MASK, k[1] != 0ull));

  // 4. greedy NMS, tile by tile
  const float thr = P.iou_threshold;
  int nk = 0;
  u64 last_key = 0ull;   // key of the last box kept so far (uniform)
  for (int tile = 0; tile < 2 && nk < P.M_cap; ++tile) {
    const u64 key = k[tile];
    bool alive = key != 0ull;
    const u32 cand_any = __ballot_sync(RPP_FULL_MASK, alive);
    if (cand_any == 0u) break;
    float4 orig = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bx = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    float area = 0.0f;
    if (alive) {
      orig = col_box<false>(P, b, c, key_tie(key));
      if (P.clip_before) orig = clip01(orig);
      const float4 cb = canon_box(orig, area);
      if (area > 0.0f) bx = cb; else area = 0.0f;
    }
    for (int q = 0; q < nk && alive; ++q)
      if (iou_gt(bx, area, sh->kbox[q], sh->karea[q], thr)) alive = false;
    // The probe keeps only M_cap (a handful of) boxes: walk the survivors in order — the best remaining candidate is
    // kept, its box is broadcast, every later candidate tests itself against it — instead of building the full
    // 32 x 32 suppression mask first (one IoU per lane per kept box instead of 31 per lane).
    u32 alive_bits = __ballot_sync(RPP_FULL_MASK, alive);
    while (alive_bits != 0u && nk < P.M_cap) {
      const int i = __ffs(alive_bits) - 1;
      const float4 kb = make_float4(__shfl_sync(RPP_FULL_MASK, bx.x, i), __shfl_sync(RPP_FULL_MASK, bx.y, i),
                                    __shfl_sync(RPP_FULL_MASK, bx.z, i), __shfl_sync(RPP_FULL_MASK, bx.w, i));
      const float ka = __shfl_sync(RPP_FULL_MASK, area, i);
      if (lane == i) {
        sh->kbox[nk] = bx;
        sh->karea[nk] = area;
        P.sel_key[p * P.M + nk] = key;
        P.sel_box[p * P.M + nk] = orig;
      }
      last_key = __shfl_sync(RPP_FULL_MASK, key, i);
      ++nk;
      alive_bits &= ~(1u << i);
      const bool sup = ((alive_bits >> lane) & 1u) && iou_gt(bx, area, kb, ka, thr);
      alive_bits &= ~__ballot_sync(RPP_FULL_MASK, sup);
    }
    __syncwarp();
  }
  if (lane == 0) {
    P.sel_cnt[p] = nk;
    float bd;
    if (nk >= P.M_cap && nk > 0) bd = key_score(last_key);
    else if ((long)n_valid >= P.k_lim) bd = -INFINITY;                         // consumed all the class may consume
    else if (whole_list && list_complete) bd = -INFINITY;                      // class exhausted
    else bd = whole_list ? s_edge : e0;                                        // the rest scores <= this
    P.bound[p] = bd;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Top-k emission fast path (FilterTopKDetections, rpp_topk and the global pre-NMS filter): when a problem's whole
// candidate list fits in shared memory, one 1024-thread block scores it, sorts it once (bitonic, <= 16 K keys) and
// writes the k best keys.  Problems it cannot serve exactly (list overflowed, too long, or fewer than k candidates
// safely above the collect edge) are left to the generic lazy kernel, which skips the ones done here.
// ---------------------------------------------------------------------------------------------------------------
#define RPP_EMIT_NT 1024
#define RPP_EMIT_CAP 16384
#define RPP_EMIT_CHUNK 8192
struct EmitShared {
  SelectScratch<RPP_EMIT_NT> sel;
  u64 keys[RPP_EMIT_CAP];
  u64 chunk[RPP_EMIT_CHUNK];
  int valid;
};
// Scored keys of one emission problem in sh->keys[0 .. n_keys): the candidate list (or, when it cannot serve k rows,
// an in-block exact re-collection of the column).  Returns the number of VALID keys (strictly above everything that is
// not in sh->keys) — at least k_lim — or -1 when the block cannot serve the problem (the generic kernel takes it).
__device__ __forceinline__ int emit_prepare(const ColProblemParams& P, EmitShared* sh, size_t p, int b, int c,
                                            int& n_keys) {
  const int tid = threadIdx.x;
  if (tid == 0) sh->valid = 0;
  n_keys = 0;       // scored keys in sh->keys[0 .. n_keys)
  const u32 n_raw = P.cand_count[p];
  if (n_raw & 0x80000000u) return -1;
  const bool list_ok = !(P.force_scan || n_raw > (u32)P.CAP || n_raw > RPP_EMIT_CAP || n_raw == 0);
  int nv = 0;       // of which valid
  __syncthreads();
  if (list_ok) {
    n_keys = (int)n_raw;
    const float T = P.T[p];
    const bool list_complete = !(T > P.T_min);
    const float s_edge = list_complete ? P.score_threshold : col_score(P, T);
    const uint2* lst = P.cand + p * (size_t)P.CAP;
    int local = 0;
    for (int i = tid; i < n_keys; i += RPP_EMIT_NT) {
      const uint2 e = lst[i];
      const float sc = col_score(P, __uint_as_float(e.x));
      u64 k = 0ull;
      if (sc > P.score_threshold && (list_complete || sc > s_edge)) { k = make_key(sc, e.y); ++local; }
      sh->keys[i] = k;
    }
    __syncthreads();
    if (local) atomicAdd(&sh->valid, local);
    __syncthreads();
    nv = sh->valid;
  }
  if ((long)nv < P.k_lim) {
    // The list came up short of k (the sampled threshold was too high), overflowed or does not exist.  Instead of
    // leaving the problem to repeated scored scans of the column (tens of milliseconds on the flat 6 M element axis
    // of the global filter), collect again INSIDE the block with an exact cut: one radix select over the column's
    // RAW keys (value bits | ~row; no sigmoid) delivers its best ~1.1 k .. 16 K rows, everything else is below the
    // cut; the selected rows are scored and, by the edge rule, those strictly above the score of the cut are complete.
    if (P.force_scan) return -1;   // debug: the generic kernel's exact scan is what is being tested
    __syncthreads();
    u64 KBr = ~0ull;
    u32 population = 0u;
    long want = P.k_lim + P.k_lim / 8 + 64;
    if (want > RPP_EMIT_CAP) return -1;   // more than one block's worth: the generic kernel takes it
    if (tid == 0 && P.scan_count) atomicAdd(P.scan_count, 1u);
    const int m = select_chunk<RPP_EMIT_NT>(
        [&](int i) -> u64 {
          const float raw = lv_val(P.lv, b, i, P.C, c);
          return raw >= P.T_min ? (((u64)ord_f32(raw) << 32) | (u64)(0xffffffffu - (u32)i)) : 0ull;
        },
        (int)P.N, KBr, (int)want, sh->keys, RPP_EMIT_CAP, &sh->sel, /*sort=*/false, &population);
    const bool whole = (u32)m == population;        // every eligible row of the column was selected
    const float e0 = whole ? P.score_threshold : col_score(P, unord_f32((u32)(KBr >> 32)));
    if (tid == 0) sh->valid = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < m; i += RPP_EMIT_NT) {
      const u64 rk = sh->keys[i];
      const float sc = col_score(P, unord_f32((u32)(rk >> 32)));
      u64 k = 0ull;
      if (sc > P.score_threshold && (whole || sc > e0)) { k = make_key(sc, key_tie(rk)); ++local; }
      sh->keys[i] = k;
    }
    __syncthreads();
    if (local) atomicAdd(&sh->valid, local);
    __syncthreads();
    n_keys = m;
    nv = sh->valid;
    if ((long)nv < P.k_lim) return -1;   // a huge tie group at the cut, or a coarse radix cut: the generic kernel decides
  }
  return nv;
}

// RAW keys (ordered logit bits | ~row) of one emission problem in sh->keys[0 .. n): the candidate list when it holds at
// least k_lim elements, else an in-block exact re-collection of the column's best rows (as in emit_prepare).  The keys
// are ALL elements of the column whose ordered logit encoding is >= o_complete.  Returns n >= k_lim, or -1 when the
// block cannot serve the problem.
// kmin / kmax: per-THREAD extrema of the keys the thread handled (~0 / 0 when it handled none): the caller reduces them.
__device__ __forceinline__ int emit_prepare_raw(const ColProblemParams& P, EmitShared* sh, size_t p, int b, int c,
                                                u32& o_complete, u64& kmin, u64& kmax) {
  const int tid = threadIdx.x;
  o_complete = 0xffffffffu;
  kmin = ~0ull; kmax = 0ull;
  const u32 n_raw = P.cand_count[p];
  const float T = P.T[p];        // (both loads in flight together)
  if (n_raw & 0x80000000u) return -1;
  if (P.force_scan) return -1;   // debug: the generic kernel's exact scan is what is being tested
  const bool list_ok = !(n_raw > (u32)P.CAP || n_raw > RPP_EMIT_CAP || (long)n_raw < P.k_lim);
  if (list_ok) {
    const uint2* lst = P.cand + p * (size_t)P.CAP;
#pragma unroll 4
    for (int i = tid; i < (int)n_raw; i += RPP_EMIT_NT) {
      const uint2 e = lst[i];
      const u64 k = ((u64)ord_f32(__uint_as_float(e.x)) << 32) | (u64)(0xffffffffu - e.y);
      sh->keys[i] = k;
      kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax;
    }
    __syncthreads();
    o_complete = ord_f32(T > P.T_min ? T : P.T_min);
    return (int)n_raw;
  }
  u64 KBr = ~0ull;
  u32 population = 0u;
  const long want = P.k_lim + P.k_lim / 8 + 64;
  if (want > RPP_EMIT_CAP) return -1;   // more than one block's worth: the generic kernel takes it
  if (tid == 0 && P.scan_count) atomicAdd(P.scan_count, 1u);
  const int m = select_chunk<RPP_EMIT_NT>(
      [&](int i) -> u64 {
        const float raw = lv_val(P.lv, b, i, P.C, c);
        return raw >= P.T_min ? (((u64)ord_f32(raw) << 32) | (u64)(0xffffffffu - (u32)i)) : 0ull;
      },
      (int)P.N, KBr, (int)want, sh->keys, RPP_EMIT_CAP, &sh->sel, /*sort=*/false, &population);
  if ((long)m < P.k_lim) return -1;
  for (int i = tid; i < m; i += RPP_EMIT_NT) {
    const u64 k = sh->keys[i];
    kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax;
  }
  // rows with the cut's own logit may have been left out (higher row index): complete strictly above it
  o_complete = (u32)m == population ? ord_f32(P.T_min) : (u32)(KBr >> 32) + 1u;
  return m;
}

#define RPP_EMIT_BIN_MAX 256   // largest bin the bucket sort ranks by counting
// Writes the k_lim best of the scored keys sh->keys[0 .. n_keys) (0 = not a key; at least k_lim are) to emit_key in
// descending order.  Returns false (nothing written that matters) when the key distribution does not suit it.
__device__ bool emit_bucket_sort(const ColProblemParams& P, EmitShared* sh, size_t p, int n_keys) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u32* hist = sh->sel.hist;     // [1024] counts, then scatter cursors
  u32* offs = sh->sel.part;     // [1024] keys in the bins above (thread t owns bin t)
  __shared__ u32 s_dk, s_total, s_bad;
  u32 cnt = 0;
  u64 mx = 0ull, mn = ~0ull;
  for (int i = tid; i < n_keys; i += RPP_EMIT_NT) {
    const u64 k = sh->keys[i];
    if (k != 0ull) { ++cnt; mx = k > mx ? k : mx; mn = k < mn ? k : mn; }
  }
  block_cnt_max_min<RPP_EMIT_NT>(cnt, mx, mn, &sh->sel.bs);
  if ((long)cnt < P.k_lim) return false;
  const u64 q = (mx - mn) / 1024ull + 1ull;          // bin = floor((key - mn) / q) < 1024 (any monotone map would do)
  const u64 rq = q > 1ull ? ~0ull / q : 0ull;        // umulhi(x, floor((2^64 - 1) / q)) <= x / q, monotone in x
  auto bin_of = [&](u64 k) -> u32 { return q > 1ull ? (u32)__umul64hi(k - mn, rq) : (u32)(k - mn); };
  hist[tid] = 0u;
  if (tid == 0) s_bad = 0u;
  __syncthreads();
  for (int i = tid; i < n_keys; i += RPP_EMIT_NT) {
    const u64 k = sh->keys[i];
    if (k != 0ull) atomicAdd(&hist[bin_of(k)], 1u);
  }
  __syncthreads();
  {
    const u32 h = hist[tid];
    u32 v = h;   // inclusive suffix count over the bins >= tid
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_down_sync(RPP_FULL_MASK, v, o);
      if (lane + o < 32) v += t;
    }
    __shared__ u32 s_wtot[32], s_wsfx[32];
    if (lane == 0) s_wtot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      const u32 tot = s_wtot[lane];
      u32 sfx = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_down_sync(RPP_FULL_MASK, sfx, o);
        if (lane + o < 32) sfx += t;
      }
      s_wsfx[lane] = sfx - tot;
    }
    __syncthreads();
    const u32 above = v - h + s_wsfx[warp];
    offs[tid] = above;
    const u32 kth = (u32)P.k_lim;
    if (above < kth && above + h >= kth) { s_dk = (u32)tid; s_total = above + h; }
    if (above < kth && h > RPP_EMIT_BIN_MAX) s_bad = 1u;   // a crowded bin at or above the k-th
  }
  __syncthreads();
  const u32 d_k = s_dk, total = s_total;
  if (s_bad || total > RPP_EMIT_CHUNK) return false;
  hist[tid] = 0u;   // cursors
  __syncthreads();
  for (int i = tid; i < n_keys; i += RPP_EMIT_NT) {
    const u64 k = sh->keys[i];
    if (k == 0ull) continue;
    const u32 bin = bin_of(k);
    if (bin >= d_k) sh->chunk[offs[bin] + atomicAdd(&hist[bin], 1u)] = k;
  }
  __syncthreads();
  u64* out = P.emit_key + p * (size_t)P.k_lim;
  for (u32 i = tid; i < total; i += RPP_EMIT_NT) {
    const u64 k = sh->chunk[i];
    const u32 bin = bin_of(k);
    const u32 s0 = offs[bin], e0 = s0 + hist[bin];
    u32 rank = s0;
    for (u32 j = s0; j < e0; ++j) rank += sh->chunk[j] > k;
    if ((long)rank < P.k_lim) out[rank] = k;
  }
  return true;
}

__device__ __forceinline__ void emit_sort_body(const ColProblemParams& P, EmitShared* sh, const size_t p) {
  const int tid = threadIdx.x;
  const int b = (int)(p / P.C), c = (int)(p % P.C);
  if (P.emit_direct && P.emit_done[p] == 2) return;   // global_top_direct_kernel already wrote the image's detections
  if (tid == 0) P.emit_done[p] = 0;
  int n_keys = 0;
  if (emit_prepare(P, sh, p, b, c, n_keys) < 0) return;
  // ---- bucket sort: ONE histogram over a linear map of [min key, max key] onto 1024 bins orders the keys up to their
  // bin (~10 keys per bin for a list of 10^4); the bins from the top down to the one that holds the k-th key are
  // scattered into place and ranked inside themselves by counting.  Replaces a radix cut (whose MSB-first digits put a
  // whole list into a dozen bins: the scores of a list share their exponent) plus an 8 192-key bitonic sort; lists with
  // a crowded bin (heavy ties) or more than a chunk's worth above the k-th bin take that path below.
  if (emit_bucket_sort(P, sh, p, n_keys)) {
    if (tid == 0) P.emit_done[p] = 1;
    return;
  }
  __syncthreads();
  // the k best keys, in order: usually ONE exact radix cut to [k, 8192] keys and one bitonic sort of that chunk
  u64 KB = ~0ull;
  long emitted = 0;
  while (emitted < P.k_lim) {
    const long want = P.k_lim - emitted;
    const int m = select_chunk<RPP_EMIT_NT>([&](int i) { return sh->keys[i]; }, n_keys, KB,
                                            (int)(want < RPP_EMIT_CHUNK ? want : RPP_EMIT_CHUNK), sh->chunk,
                                            RPP_EMIT_CHUNK, &sh->sel);
    if (m == 0) break;
    const long take = (long)m < want ? m : want;
    for (long i = tid; i < take; i += RPP_EMIT_NT) P.emit_key[p * (size_t)P.k_lim + emitted + i] = sh->chunk[i];
    emitted += take;
    __syncthreads();
  }
  if (tid == 0 && emitted == P.k_lim) P.emit_done[p] = 1;
}

__global__ void __launch_bounds__(RPP_EMIT_NT) emit_sort_kernel(ColProblemParams P) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EmitShared* sh = reinterpret_cast<EmitShared*>(smem_raw);
  if (P.n_loop > 0) {
    for (size_t p = blockIdx.x; (long)p < P.n_loop; p += gridDim.x) {
      emit_sort_body(P, sh, p);
      __syncthreads();
    }
  } else {
    emit_sort_body(P, sh, blockIdx.x);
  }
}
