// rpp_api.cu — C ABI of libretinapost.so (include/retinapost.h): handle, workspace layout and kernel launches.
// Host logic only; every result is computed by the kernels in rpp_kernels.cuh.  There is no CPU compute path.
#include <algorithm>
#include <numeric>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "../../include/retinapost.h"
#include "rpp_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail(RPP_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__));      \
  } while (0)

#define LAUNCHED()                                                                                 \
  do {                                                                                             \
    ++g_launches;                                                                                  \
    cudaError_t e__ = cudaGetLastError();   /* (clears it: a failed launch must not poison later checks) */ \
    if (e__ != cudaSuccess) return fail(RPP_ECUDA, "kernel launch (%s:%d): %s", __FILE__, __LINE__, \
                                        cudaGetErrorString(e__));                                  \
  } while (0)

// Candidate list capacity per problem of a sampled plan (env RPP_LIST_CAP, default 8192): a column whose population
// above the score threshold is at most ~0.7 of it is collected whole (see make_plan's complete-list rule).
static int kListCap = [] {
  const char* v = getenv("RPP_LIST_CAP");
  const int c = v ? atoi(v) : 8192;
  return c >= 1024 && c <= 65536 ? c : 8192;
}();

struct SamplePlan {
  bool on;
  int stride, G, lanes, rows_per_group, rank;
  int rank_lo;   // complete-list rule of the rank kernels (-1: off)
  int CAP;
};

struct Handle {
  rpp_config cfg;
  std::vector<double> areas, ratios, scales;
  long N;
  int levels;
  AnchorParams ap;
  float4* d_anchors;
  u32* d_scan_count;     // problems that left their candidate list for an exact scan of the column (rpp_debug_exact_scans)
  float T_logit;   // smallest logit whose sigmoid exceeds score_threshold
  int device;
  int sm_count;
  int force_scan;
  int warp_probe;        // env RPP_WARP_PROBE (default 1): warp-per-problem probe kernel for the hard modes
  int probe_extra;       // env RPP_PROBE_EXTRA: boxes per class kept by the probe beyond ceil(M / C)
  int finish_argmax;     // env RPP_FINISH_ARGMAX (default 1): argmax-iterate consumer in the finish pass of the hard modes
  int top_direct;        // env RPP_TOP_DIRECT (default 1): GlobalHardNMS behind the global filter straight from the lists
  int pdl;               // env RPP_PDL (default 1): programmatic dependent launch between the kernels of a pipeline
  int two_pass;          // env RPP_TWO_PASS (default 1): probe / bound / finish scheme of the per-class modes
  int collect_ctas;      // env RPP_COLLECT_CTAS: CTAs per SM of the collect kernel (0 = automatic)
  int overlap_hint;
  int overlap;           // run NMS of image chunk i on a side stream under the collect stream of chunk i+1
  int target;            // candidates per problem the sampled pre-threshold aims at (env RPP_TARGET)
  cudaStream_t side;
  cudaEvent_t ev_chunk[8];
  cudaEvent_t ev_join;
  long tile_rows;        // tuning knob (env RPP_TILE_ROWS): anchors per collect tile, 0 = automatic
  int emit_short;        // test knob (env RPP_EMIT_SHORT): aim the top-k lists at k/2 so that every problem falls back
  int half_variant;      // tuning knob (env RPP_HALF_VARIANT): 16-bit collect, 0 = unroll 4 / 3 CTAs per SM, 1 = 8 / 2
  int collect_variant;   // tuning knob (env RPP_COLLECT_VARIANT): 0 = unroll 4 / 3 CTAs per SM, 1 = 4/2, 2 = 8/2
  DecodeParams dp;
  // optional per-stage timing (bench.py roofline): 5 events per call = boundaries of sample|collect|nms|merge
  int timing;
  std::vector<cudaEvent_t> events;       // pool, ev_used of them recorded since timing was switched on
  std::vector<const char*> ev_label;     // label of the segment that ENDS at the event (nullptr: a chain starts)
  size_t ev_used;
  int timed_calls;
  // rpp_detect_host staging (allocated on first use)
  struct HostPath {
    int chunk = 0, B_out = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    float* d_logits[2] = {nullptr, nullptr};
    float* d_deltas[2] = {nullptr, nullptr};
    void* ws = nullptr;
    size_t ws_bytes = 0;
    float* d_boxes = nullptr; float* d_scores = nullptr; void* d_classes = nullptr; int* d_valid = nullptr;
  } hp;
};

const int kStages = 4;
const size_t kMaxTimedEvents = 1 << 16;

// Per-stage timing (rpp_debug_stage_timing): every pipeline drops named marks on its stream; a mark closes the segment
// that started at the previous mark of the same call chain (label nullptr opens a chain).  Labels are static strings
// "bucket:kernel(s)" with bucket in {sample, collect, nms, merge, emit, rows}.
inline void stage_mark(Handle* h, const char* label, cudaStream_t st) {
  if (!h->timing || h->ev_used >= kMaxTimedEvents) return;
  while (h->events.size() <= h->ev_used) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    h->events.push_back(e);
    h->ev_label.push_back(nullptr);
  }
  cudaEventRecord(h->events[h->ev_used], st);
  h->ev_label[h->ev_used++] = label;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Kernel launch through cudaLaunchKernelEx, optionally with programmatic dependent launch (see pdl_enter in
// rpp_common.cuh): used for every kernel of a pipeline that directly follows another kernel in the stream.
template <class... Params, class... Args>
inline void launch_k(bool pdl, void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}


// Pre-threshold plan for columns of n rows and C classes: aim at ~`target` candidates per problem with list
// capacity 4096; small columns are collected whole (no sampling).
SamplePlan make_plan(long n, int C, int target, int lane_cap = 12, double q_plan = 0.0) {
  SamplePlan s{};
  s.on = false;
  s.rank_lo = -1;
  s.CAP = (int)n;
  if (n < 16384 || n <= 8L * target || C > 1024) return s;
  int lanes = 1024 / C;
  if (lanes > lane_cap) lanes = lane_cap;
  const int G = lanes * RPP_GPT;
  // group size chosen so that the wanted logit sits near the 70th percentile of the group maxima:
  // rows per group g = -ln(0.7) * n / target, i.e. one sampled row every S = target / (0.357 * G) rows (~3 %)
  static const double neg_ln_q = [] {   // tuning knob: the quantile of the group maxima the plan aims at
    const char* v = getenv("RPP_SAMPLE_Q");
    const double q = v ? atof(v) : 0.7;
    return -std::log(q > 0.05 && q < 0.97 ? q : 0.7);
  }();
  // (q_plan: the fine emission plans have 4x the groups and can afford a sparser sample: quantile 0.8, 1.6x fewer loads,
  // the list length's 1-sigma goes from 9 % to 11 % with both tails still > 4 sigma away)
  const double nlq = q_plan > 0.0 ? -std::log(q_plan) : neg_ln_q;
  int S = (int)std::floor(target / (nlq * G));
  if (S < 1) S = 1;
  long g = n / ((long)S * G);
  if (g < 8) {   // the target is a large fraction of the column: sample denser, accept a lower quantile
    g = 8;
    S = (int)(n / (g * G));
    if (S < 1) return s;
  }
  const double q = std::pow(1.0 - (double)target / (double)n, (double)g);
  if (q < 0.05 || q > 0.97 || G < 8) return s;
  int rank = (int)std::floor(q * G);
  if (rank < 2) rank = 2;
  if (rank > G - 3) rank = G - 3;
  s.on = true;
  s.stride = S;
  s.G = G;
  s.lanes = lanes;
  s.rows_per_group = (int)g;
  s.rank = rank;
  s.CAP = kListCap;
  // complete-list rule: a column with at most ~0.7 * CAP elements above the score threshold is collected whole
  // (threshold = T_min).  Such a column leaves a fraction q0 = (1 - 0.7 CAP / n)^g of the groups without any element
  // above T_min; with a true population of CAP the expected fraction is far lower (> 3 sigma at G = 96), and an
  // overflowing list is still handled exactly (exact scan), just slowly.
  {
    const double q0 = std::pow(1.0 - 0.7 * s.CAP / (double)n, (double)g);
    s.rank_lo = (n > 2L * s.CAP && q0 > 0.02) ? (int)std::floor(q0 * G) : -1;
    if (s.rank_lo >= G - 1) s.rank_lo = G - 2;
  }
  return s;
}

const int kTarget = 768;

// The plan of one problem set: NMS problems aim at `nms_target` candidates per column; top-k emission (k_lim rows must
// come out of the list) aims well above k_lim — see the comment in run_problem_set.  Pure host arithmetic (exported
// for the CPU tests as rpp_debug_sample_plan).
SamplePlan choose_plan(long n, int C, bool emit, long k_lim, int nms_target, int emit_short, bool* emit_fine_out,
                       int* target_out) {
  const int emit_lanes = std::min(1024 / std::max(1, C), 48);
  const bool emit_fine = emit && emit_lanes * RPP_GPT >= 256 && n >= (1L << 19);   // short columns: the scan is cheap
  const int target = emit ? (int)std::min<long>(emit_fine ? 2 * k_lim + 256 : k_lim + k_lim / 2 + 512, 1 << 28)
                          : nms_target;
  SamplePlan plan = make_plan(n, C, emit_short && emit ? (int)std::max<long>(64, k_lim / 2) : target,
                              emit_fine ? emit_lanes : 12, emit_fine ? 0.8 : 0.0);
  if (plan.on && emit) { plan.CAP = emit_fine ? target + target / 2 + 1024 : 2 * target; plan.rank_lo = -1; }
  if (emit_fine_out) *emit_fine_out = emit_fine;
  if (target_out) *target_out = target;
  return plan;
}

// Bump allocator over the caller's workspace.  Every pipeline below is written once and run twice: a dry pass
// (no launches) sizes the workspace — rpp_workspace_bytes and the capacity check share it — then the real pass.
struct Arena {
  char* base;
  size_t off;
  bool dry;
  template <class T> T* take(size_t count) {
    const size_t o = off;
    off = align_up(off + count * sizeof(T), 256);
    return dry ? nullptr : reinterpret_cast<T*>(base + o);
  }
};

// rpp_config.tpu_semantics: the TPUStrategy branches of GenerateDetections replace the two hard modes
// (postprocessing_ops.py:549-550, :558-559); rpp_create rejects the flag with any other mode, as the reference's
// constructor does under a TPUStrategy (:202-206).
bool tpu_branch(const rpp_config& c) {
  return c.tpu_semantics && (c.mode == RPP_GLOBAL_HARD_NMS || c.mode == RPP_PER_CLASS_HARD_NMS);
}

bool is_per_class_mode(int mode) {
  return mode == RPP_COMBINED_NMS || mode == RPP_PER_CLASS_HARD_NMS || mode == RPP_PER_CLASS_SOFT_NMS;
}

// One set of P = B * C "column problems": columns of x [B, n, C] -> lazily sorted candidates -> consumer.
struct ProblemSet {
  // in
  const float* x; int is_logit; int B; long n; int C;   // fused tensors ...
  const float4* deltas; const float4* boxes; int q;
  const Levels* levels;    // ... or the per-level pieces (x / deltas then unused)
  const u64* row_keys; long k_rows; int C_src; const Levels* delta_lv;   // rows resolved through sorted keys (Global*)
  int consumer;            // RPP_CONSUME_*
  long k_lim; int M_lim; int M;
  int clip_before; float iou_threshold; float score_threshold; float T_min;
  float soft_sigma_tf; int tie_is_rank;
  int two_pass_m1;         // > 0: probe with this many kept per class, bound per image, finish (per-class modes)
  int padded;              // RPP_CONSUME_PADDED: 1 global (score filter inside), 2 per class (no score filter inside)
  int row0_mode;
  const GlobalTopDirectParams* direct;   // emission for GlobalHardNMS behind the global filter: global_top_direct_kernel first
  // out (workspace)
  int* emit_done;
  u64* sel_key; float4* sel_box; int* sel_cnt; u64* emit_key;
  float* pad_score; float4* pad_box;
};

int run_problem_set(Handle* h, Arena& ar, ProblemSet& ps, cudaStream_t st, cudaStream_t st2 = nullptr,
                    cudaEvent_t ev = nullptr) {
  const int B = ps.B, C = ps.C;
  const long n = ps.n;
  const size_t P = (size_t)B * C;
  const bool emit = ps.consumer == RPP_CONSUME_EMIT;
  // emission must reach k_lim from the list (falling short means an exact scan of the whole column, which costs
  // tens of milliseconds on a 6 M element column): the estimate's 1-sigma error is ~17 % with 96 group maxima, so
  // narrow problem sets (few classes: the global filter, C = 1) sample 4x as many groups (1-sigma ~8 %) and aim at
  // 2 k candidates — both tails (fewer than k, more than the list capacity) are then > 5 sigma away
  bool emit_fine = false;
  int target = 0;
  SamplePlan plan = choose_plan(n, C, emit, ps.k_lim, h->target, h->emit_short, &emit_fine, &target);
  const size_t gm_elems = plan.on ? (size_t)B * plan.G * C : 0;

  float* T = ar.take<float>(P);
  u32* cand_count = ar.take<u32>(P);
  u32* tile_counter = ar.take<u32>(64);
  u32* gm = ar.take<u32>(gm_elems);
  const size_t zero_bytes = ar.dry ? 0 : (size_t)((char*)gm - (char*)cand_count) + gm_elems * sizeof(u32);
  uint2* cand = ar.take<uint2>(P * (size_t)plan.CAP);
  ps.sel_cnt = nullptr; ps.sel_key = nullptr; ps.sel_box = nullptr; ps.emit_key = nullptr;
  u64* r_key = nullptr; uint2* r_meta = nullptr; float4* r_box = nullptr;
  long r_cap = 0;
  int* emit_done = nullptr;
  if (emit) {
    ps.emit_key = ar.take<u64>(P * (size_t)ps.k_lim);
    emit_done = ar.take<int>(P);
    ps.emit_done = emit_done;
  } else {
    ps.sel_cnt = ar.take<int>(P);
    ps.sel_key = ar.take<u64>(P * (size_t)ps.M);
    ps.sel_box = ar.take<float4>(P * (size_t)ps.M);
    if (ps.consumer == RPP_CONSUME_SOFT) {
      r_cap = std::max<long>(0, std::min<long>(ps.k_lim, n) - RPP_SOFT_RS) + 1;
      r_key = ar.take<u64>(P * (size_t)r_cap);
      r_meta = ar.take<uint2>(P * (size_t)r_cap);
      r_box = ar.take<float4>(P * (size_t)r_cap);
    }
  }
  ps.pad_score = nullptr; ps.pad_box = nullptr;
  if (ps.padded == 2) {
    ps.pad_score = ar.take<float>(P);
    ps.pad_box = ar.take<float4>(P);
  }
  float* bound = nullptr;
  float* stop_L = nullptr;
  u32* work_items = nullptr;
  if (ps.two_pass_m1 > 0) {
    bound = ar.take<float>(P);
    stop_L = ar.take<float>((size_t)B);
    work_items = ar.take<u32>(P);
  }
  if (ar.dry) return RPP_OK;

  Levels lv;
  memset(&lv, 0, sizeof(lv));
  if (ps.levels) lv = *ps.levels;
  else { lv.L = 1; lv.off[0] = 0; lv.off[1] = n; lv.x[0] = ps.x; lv.d[0] = ps.deltas; }
  bool aligned = true;
  for (int l = 0; l < lv.L; ++l) aligned = aligned && ((uintptr_t)lv.x[l] % 16) == 0;

  // ---- stage 0/1: thresholds ---------------------------------------------------------------------------------
  CUDA_OK(cudaMemsetAsync(cand_count, 0, zero_bytes, st));
  stage_mark(h, nullptr, st);
  if (plan.on && !h->force_scan) {
    const int threads = (int)align_up((size_t)plan.lanes * C, 32);
    // narrow problem sets (C = 1: 48 active threads per block) need more blocks in flight to cover the load latency
    static const int split_env = [] { const char* v = getenv("RPP_SAMPLE_SPLIT"); return v ? atoi(v) : 0; }();
    const int max_split = split_env > 0 ? split_env : (threads <= 64 ? 64 : 16);
    int split = plan.rows_per_group < max_split ? plan.rows_per_group : max_split;
    // two 960-thread blocks per SM: about two full waves of blocks (configs[1]: 64 x 9 = 576 of 592) instead of 3.5
    // (64 x 16) — the tail of the last partial wave cost 2 us of the 23
    if (split_env <= 0 && threads > 64 && (long)B * split > 4L * h->sm_count)
      split = std::max(1, (int)(4L * h->sm_count / B));
    if (C == 1 && lv.L == 1 && lv.dtype == RPP_DT_F32 && plan.rows_per_group >= 8) {
      const int lead = (int)(((uintptr_t)lv.x[0] % 16) / 4);
      sample_max_flat4_kernel<<<dim3(B, std::min(plan.rows_per_group / 4, max_split)), threads, 0, st>>>(
          lv.x[0] - lead, lead, n, plan.stride, plan.lanes, plan.rows_per_group / 4, gm);
    }
    else if (lv.L > 1 && lv.dtype != RPP_DT_F32)
      sample_max_kernel<true, true><<<dim3(B, split), threads, 0, st>>>(lv, n, C, plan.stride, plan.lanes,
                                                                        plan.rows_per_group, gm);
    else if (lv.L > 1)
      sample_max_kernel<true, false><<<dim3(B, split), threads, 0, st>>>(lv, n, C, plan.stride, plan.lanes,
                                                                         plan.rows_per_group, gm);
    else if (lv.dtype != RPP_DT_F32)
      sample_max_kernel<false, true><<<dim3(B, split), threads, 0, st>>>(lv, n, C, plan.stride, plan.lanes,
                                                                         plan.rows_per_group, gm);
    else
      sample_max_kernel<false, false><<<dim3(B, split), threads, 0, st>>>(lv, n, C, plan.stride, plan.lanes,
                                                                          plan.rows_per_group, gm);
    LAUNCHED();
    if (plan.G <= 128) {
      launch_k(h->pdl, sample_rank_sort_kernel, dim3(dim3(B, (C + RPP_RANK_CPB - 1) / RPP_RANK_CPB)), dim3(RPP_RANK_CPB * 32), 0, st, 
          gm, C, plan.G, plan.rank, plan.rank_lo, ps.T_min, T);
    } else {
      const size_t smem = (size_t)plan.G * RPP_RANK_CPB * sizeof(u32);
      launch_k(h->pdl, sample_rank_kernel, dim3(dim3(B, (C + RPP_RANK_CPB - 1) / RPP_RANK_CPB)), dim3(256), smem, st, gm, C, plan.G, plan.rank,
                                                                                          plan.rank_lo, ps.T_min, T);
    }
    LAUNCHED();
  } else {
    fill_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(T, P, ps.T_min);
    LAUNCHED();
  }
  stage_mark(h, emit ? "emit:sample" : "sample", st);
  // ---- stage 2: collect --------------------------------------------------------------------------------------
  // Unsampled SCORE columns (dense stage inputs of a few thousand rows): a list would just be a copy of the column,
  // so the problem kernel scans the column itself (its "exact scan" phase costs no sigmoid on scores).
  const bool scan_only = h->force_scan || (!plan.on && !ps.is_logit && !emit);
  // the vectorised column collects stage RPP_STAGE_CAP hits per class in shared memory: ~520 bytes per class
  const bool stage_fits = (size_t)C * RPP_STAGE_CAP * sizeof(uint2) + 2 * (size_t)C * sizeof(u32) <= 200 * 1024;
  if (!scan_only) {
    if (C == 1) {
      // single column (the flat anchors x classes axis, or row maxima), any n and any element alignment, fp32 / f16 /
      // bf16, fused or in per-level pieces: each piece is one flat array of 128-bit words (collect_flat_kernel)
      for (int l = 0; l < lv.L; ++l) {
        const long n_l = lv.off[l + 1] - lv.off[l];
        const int es = lv.dtype == RPP_DT_F32 ? 4 : 2, EPW = 16 / es;
        const int lead = (int)(((uintptr_t)lv.x[l] % 16) / es);
        const char* xa = reinterpret_cast<const char*>(lv.x[l]) - (size_t)lead * es;
        const int UNROLL = 4;
        const long unit = (long)RPP_COLLECT_NT * UNROLL * EPW;   // elements per block-wide load round
        long tile_elems = 4 * unit;
        if (plan.on) {   // ~a third of the shared queue (RPP_FLAT_QCAP) expected per tile
          const long want = (long)((RPP_FLAT_QCAP / 3.0) * (double)n / target);
          tile_elems = std::max(unit, std::min(want / unit * unit, 16 * unit));
        }
        const long total = (long)lead + (long)B * n_l;
        {   // small batches: at least ~4 tiles per resident CTA
          const long want_tiles = 4L * h->sm_count * 3;
          while (tile_elems > unit && (total + tile_elems - 1) / tile_elems < want_tiles) tile_elems -= unit;
        }
        const long n_tiles = (total + tile_elems - 1) / tile_elems;
        const long grid = std::min<long>((long)h->sm_count * 3, n_tiles);
        u32* tc = tile_counter + 32 + l;   // (zeroed with the counters by the memset above)
#define RPP_LAUNCH_FLAT(DT)                                                                                      \
        launch_k(h->pdl, collect_flat_kernel<4, DT>, dim3((unsigned)grid), dim3(RPP_COLLECT_NT), 0, st,            \
                 (const void*)xa, lead, (u32)lv.off[l], T, cand_count, cand, plan.CAP, B, n_l, tile_elems, n_tiles, tc)
        if (lv.dtype == RPP_DT_F32) RPP_LAUNCH_FLAT(RPP_DT_F32);
        else if (lv.dtype == RPP_DT_F16) RPP_LAUNCH_FLAT(RPP_DT_F16);
        else RPP_LAUNCH_FLAT(RPP_DT_BF16);
#undef RPP_LAUNCH_FLAT
        LAUNCHED();
      }
    } else if (lv.dtype != RPP_DT_F32) {
      if (C % 8 != 0 || !aligned || C / 8 > RPP_COLLECT_NT || !stage_fits)
        return fail(RPP_EINVAL, "16-bit logits need num_classes % 8 == 0 (at most 392 classes) and 16-byte aligned "
                                "tensors; convert to fp32 otherwise");
      const int C8 = C / 8;
      const int lanes = RPP_COLLECT_NT / C8;
      const int UNROLL = h->half_variant == 1 ? 8 : 4;
      long rows_per_tile = plan.on ? (long)(24.0 * n / target) : 4L * lanes * UNROLL;
      rows_per_tile = (rows_per_tile + (long)lanes * UNROLL - 1) / ((long)lanes * UNROLL) * ((long)lanes * UNROLL);
      if (rows_per_tile < (long)lanes * UNROLL) rows_per_tile = (long)lanes * UNROLL;
      int tiles_per_image = 0;
      for (int l = 0; l < lv.L; ++l) {
        lv.tile_off[l] = tiles_per_image;
        tiles_per_image += (int)((lv.off[l + 1] - lv.off[l] + rows_per_tile - 1) / rows_per_tile);
      }
      lv.tile_off[lv.L] = tiles_per_image;
      const long n_tiles = (long)B * tiles_per_image;
      const size_t smem = (size_t)C * RPP_STAGE_CAP * sizeof(uint2) + 2 * (size_t)C * sizeof(u32);
#define RPP_LAUNCH_HALF(U, DT, MB)                                                                               \
      launch_k(h->pdl, collect_cols8_half_kernel<U, DT, MB>,                                                     \
               dim3((unsigned)std::min<long>((long)h->sm_count * MB, n_tiles)), dim3(RPP_COLLECT_NT), smem, st,   \
               lv, T, cand_count, cand, plan.CAP, B, n, C8, lanes, (int)rows_per_tile, tiles_per_image, tile_counter)
      if (h->half_variant == 1) {
        if (lv.dtype == RPP_DT_F16) RPP_LAUNCH_HALF(8, RPP_DT_F16, 2); else RPP_LAUNCH_HALF(8, RPP_DT_BF16, 2);
      } else {
        if (lv.dtype == RPP_DT_F16) RPP_LAUNCH_HALF(4, RPP_DT_F16, 3); else RPP_LAUNCH_HALF(4, RPP_DT_BF16, 3);
      }
#undef RPP_LAUNCH_HALF
      LAUNCHED();
    } else if (C % 4 == 0 && aligned && C / 4 <= RPP_COLLECT_NT && stage_fits) {
      const int C4 = C / 4;
      const int lanes = RPP_COLLECT_NT / C4;
      const int UNROLL = h->collect_variant == 2 ? 8 : 4;
      const int MINB = h->collect_variant == 0 ? 3 : 2;
      // tile: ~24 expected hits per class (stage capacity 64) when the plan aims at `target` candidates per column
      long rows_per_tile = plan.on ? (long)(24.0 * n / target) : 4L * lanes * UNROLL;
      rows_per_tile = (rows_per_tile + (long)lanes * UNROLL - 1) / ((long)lanes * UNROLL) * ((long)lanes * UNROLL);
      if (rows_per_tile < (long)lanes * UNROLL) rows_per_tile = (long)lanes * UNROLL;
      if (h->tile_rows > 0)
        rows_per_tile = std::max<long>((long)lanes * UNROLL, h->tile_rows / ((long)lanes * UNROLL) * ((long)lanes * UNROLL));
      {   // small batches: shrink the tiles until every resident CTA gets ~4 of them (dynamic scheduling then
          // balances the tail to within a quarter of a CTA's share)
        const long want_tiles = 4L * h->sm_count * MINB;
        const long unit = (long)lanes * UNROLL;
        while (rows_per_tile > 4 * unit && (long)B * ((n + rows_per_tile - 1) / rows_per_tile) < want_tiles)
          rows_per_tile = std::max(4 * unit, (rows_per_tile / 2 + unit - 1) / unit * unit);   // >= 4 loop trips
      }
      int tiles_per_image = 0;
      for (int l = 0; l < lv.L; ++l) {
        lv.tile_off[l] = tiles_per_image;
        tiles_per_image += (int)((lv.off[l + 1] - lv.off[l] + rows_per_tile - 1) / rows_per_tile);
      }
      lv.tile_off[lv.L] = tiles_per_image;
      const long n_tiles = (long)B * tiles_per_image;
      // resident CTAs per SM: 3 fill the register file; when NMS blocks of the previous image chunk share the SMs
      // (overlap), 2 leave them room
      int ctas = h->collect_ctas > 0 ? h->collect_ctas : ((ev || h->overlap_hint) ? 2 : MINB);
      if (ctas > MINB) ctas = MINB;
      long grid = (long)h->sm_count * ctas;
      if (grid > n_tiles) grid = n_tiles;
      const size_t smem = (size_t)C * RPP_STAGE_CAP * sizeof(uint2) + 2 * (size_t)C * sizeof(u32);
#define RPP_LAUNCH_COLLECT(U, MB)                                                                               \
      launch_k(h->pdl, collect_cols4_kernel<U, MB>, dim3((unsigned)grid), dim3(RPP_COLLECT_NT), smem, st,        \
               (const float4*)lv.x[0], T, cand_count, cand, plan.CAP, B, n, C4, lanes, (int)rows_per_tile,       \
               tiles_per_image, tile_counter)
      if (lv.L > 1)
        launch_k(h->pdl, collect_cols4_levels_kernel<4, 3>, dim3((unsigned)grid), dim3(RPP_COLLECT_NT), smem, st, 
            lv, T, cand_count, cand, plan.CAP, B, n, C4, lanes, (int)rows_per_tile, tiles_per_image, tile_counter);
      else if (h->collect_variant == 0) RPP_LAUNCH_COLLECT(4, 3);
      else if (h->collect_variant == 1) RPP_LAUNCH_COLLECT(4, 2);
      else RPP_LAUNCH_COLLECT(8, 2);
#undef RPP_LAUNCH_COLLECT
      LAUNCHED();
    } else if (lv.L > 1) {
      return fail(RPP_EINVAL, "per-level inputs need num_classes % 4 == 0 (at most 392 classes) and 16-byte aligned "
                              "level tensors; fuse them and call rpp_detect otherwise");
    } else if (C > 1 && aligned && stage_fits && C / std::__gcd(C, 4) <= RPP_COLLECT_NT && (double)n * C < 2147483647.0) {
      // any num_classes: flat 128-bit words with a class-phase-preserving stride (collect_colsv_kernel)
      const int UNROLL = 4;
      const int Cq = C / std::__gcd(C, 4);
      const int S = RPP_COLLECT_NT / Cq * Cq;
      const long unit = (long)S * UNROLL;                           // words per block-wide load round
      long rows_per_tile = plan.on ? (long)(24.0 * n / target) : 4L * unit * 4 / C + 1;
      long tile_f4 = std::max(unit, (rows_per_tile * C / 4 + unit - 1) / unit * unit);
      const long total_f4 = ((long)B * n * C) / 4;
      {
        const long want_tiles = 4L * h->sm_count * 3;
        while (tile_f4 > unit && (total_f4 + tile_f4 - 1) / tile_f4 < want_tiles) tile_f4 -= unit;
      }
      const long n_tiles = std::max<long>(1, (total_f4 + tile_f4 - 1) / tile_f4);
      const long grid = std::min<long>((long)h->sm_count * 3, n_tiles);
      const size_t smem = (size_t)C * RPP_STAGE_CAP * sizeof(uint2) + 2 * (size_t)C * sizeof(u32);
      launch_k(h->pdl, collect_colsv_kernel<4>, dim3((unsigned)grid), dim3(RPP_COLLECT_NT), smem, st, ps.x, T, cand_count, cand, plan.CAP, B, n, C, S,
                                                                         tile_f4, n_tiles, tile_counter);
      LAUNCHED();
    } else {
      const size_t tot = (size_t)B * n * C;
      size_t grid = (tot + 255) / 256;
      if (grid > (size_t)h->sm_count * 32) grid = (size_t)h->sm_count * 32;
      collect_cols1_kernel<<<(unsigned)grid, 256, 0, st>>>(ps.x, T, cand_count, cand, plan.CAP, B, n, C);
      LAUNCHED();
    }
  }
  stage_mark(h, emit ? "emit:collect" : "collect", st);
  if (ev) {  // the problems (and what follows) run on the side stream, ordered after this collect
    CUDA_OK(cudaEventRecord(ev, st));
    CUDA_OK(cudaStreamWaitEvent(st2, ev, 0));
    st = st2;
  }
  // ---- stage 3: problems -------------------------------------------------------------------------------------
  ColProblemParams pp{};
  pp.lv = lv; pp.is_logit = ps.is_logit; pp.N = n; pp.C = C;
  pp.anchors = h->d_anchors; pp.boxes = ps.boxes; pp.q = ps.q; pp.dp = h->dp;
  pp.row_keys = ps.row_keys; pp.k_rows = ps.k_rows; pp.C_src = ps.C_src;
  memset(&pp.dlv, 0, sizeof(pp.dlv));
  if (ps.delta_lv) pp.dlv = *ps.delta_lv;
  pp.clip_before = ps.clip_before;
  pp.iou_threshold = ps.iou_threshold;
  pp.score_threshold = ps.score_threshold;
  pp.T_min = ps.T_min;
  pp.M = ps.M; pp.M_lim = ps.M_lim; pp.k_lim = ps.k_lim;
  pp.T = T; pp.cand_count = cand_count; pp.cand = cand; pp.CAP = plan.CAP;
  pp.force_scan = scan_only;
  pp.scan_count = h->d_scan_count;
  pp.sel_key = ps.sel_key; pp.sel_box = ps.sel_box; pp.sel_cnt = ps.sel_cnt;
  pp.soft_scale = ps.soft_sigma_tf > 0.0f ? -0.5f / ps.soft_sigma_tf : 0.0f;
  pp.soft_ignores_iou = h->cfg.soft_ignores_iou_threshold;
  pp.tie_is_rank = ps.tie_is_rank;
  pp.r_key = r_key; pp.r_meta = r_meta; pp.r_box = r_box; pp.r_cap = r_cap;
  pp.emit_key = ps.emit_key;
  pp.emit_done = emit_done;
  pp.padded = ps.padded; pp.row0_mode = ps.row0_mode;
  pp.stop_score = ps.score_threshold;
  pp.pad_score = ps.pad_score; pp.pad_box = ps.pad_box;
  if (ps.padded == 2) {   // every row is a candidate of the NMS; the collect pass still lists the rows above the threshold
    pp.score_threshold = -INFINITY;
    pp.T_min = -INFINITY;
  }
  const size_t smem_nms = align_up(nms_shared_bytes(pp.M_lim), 16);
  unsigned emit_grid = (unsigned)P;
  auto launch = [&](void) {
    // worklist launches (finish pass): a few persistent blocks per SM instead of one block per problem
    const unsigned grid = pp.work_items ? (unsigned)std::min<size_t>(P, (size_t)h->sm_count * 4)
                                        : (pp.n_loop > 0 ? emit_grid : (unsigned)P);
    if (ps.consumer == RPP_CONSUME_HARD)
      launch_k(h->pdl, col_problem_kernel<RPP_CONSUME_HARD>, dim3(grid), dim3(RPP_NMS_NT),
               smem_nms + (pp.argmax ? RPP_LIST_SMEM * sizeof(float4) : 0), st, pp);
    else if (ps.consumer == RPP_CONSUME_PADDED)
      launch_k(h->pdl, col_problem_kernel<RPP_CONSUME_PADDED>, dim3(grid), dim3(RPP_NMS_NT), smem_nms, st, pp);
    else if (ps.consumer == RPP_CONSUME_SOFT)
      launch_k(h->pdl, col_problem_kernel<RPP_CONSUME_SOFT>, dim3(grid), dim3(RPP_NMS_NT), smem_nms + sizeof(SoftShared), st, pp);
    else
      launch_k(h->pdl, col_problem_kernel<RPP_CONSUME_EMIT>, dim3(grid), dim3(RPP_NMS_NT), smem_nms, st, pp);
  };
  pp.pass = 0; pp.M_cap = pp.M_lim; pp.want0 = 248;   // ~250 keys: a 256-wide bitonic sort
  pp.bound = bound; pp.stop_L = stop_L;
  if (ps.two_pass_m1 > 0) {
    const int m1 = ps.two_pass_m1;
    pp.pass = 1; pp.M_cap = m1; pp.want0 = std::max(24, 6 * m1);
    if (ps.consumer == RPP_CONSUME_HARD && m1 <= RPP_PROBE_MAXCAP && h->warp_probe)
      launch_k(h->pdl, probe_warp_kernel, dim3((unsigned)((P + RPP_PROBE_WARPS - 1) / RPP_PROBE_WARPS)), dim3(RPP_PROBE_WARPS * 32), 0, st, pp, P);
    else
      launch();
    LAUNCHED();
    const int bthreads = (int)std::min<size_t>(1024, align_up((size_t)C * m1 * 2, 32));   // the rank loop uses 4 lanes per box
    u32* work_ctl = tile_counter + 16;   // zeroed with the counters by the memset above
    launch_k(h->pdl, perclass_bound_kernel, dim3(B), dim3(bthreads), (size_t)C * m1 * sizeof(float), st, ps.sel_key, ps.sel_cnt, C, ps.M, m1,
                                                                              ps.M, stop_L, bound, work_items,
                                                                              work_ctl);
    LAUNCHED();
    pp.pass = 2; pp.M_cap = pp.M_lim; pp.want0 = 96;   // re-run classes usually need a few dozen boxes
    pp.argmax = ps.consumer == RPP_CONSUME_HARD && h->finish_argmax;
    pp.work_items = work_items; pp.work_ctl = work_ctl;
  }
  if (emit && ps.direct && C == 1) {   // detections straight from the lists; what follows only does the images it left
    pp.emit_direct = 1;
    // ... with a few persistent blocks walking the (almost always all done) images instead of one block per image
    pp.n_loop = (long)P;
    emit_grid = (unsigned)std::min<size_t>(P, (size_t)h->sm_count);
    launch_k(h->pdl, global_top_direct_kernel, dim3((unsigned)P), dim3(RPP_EMIT_NT), gtd_shared_bytes(), st, pp, *ps.direct);
    LAUNCHED();
    stage_mark(h, "nms:global_top_direct", st);
  }
  if (emit) {   // whole-list sort in shared memory where it applies; the generic kernel takes the rest
    launch_k(h->pdl, emit_sort_kernel, dim3(emit_grid), dim3(RPP_EMIT_NT), sizeof(EmitShared), st, pp);
    LAUNCHED();
  }
  launch();
  LAUNCHED();
  stage_mark(h, emit ? "emit:sort" : "nms", st);
  return RPP_OK;
}

struct Outputs {
  float4* boxes; float* scores; void* classes; int* valid;
};

// NonMaxSuppressionV5's (iou_threshold, soft_nms_sigma) as the reference passes them (:253-255, :448-450).
void nms_v5_args(const rpp_config& c, float* iou_thr, float* sigma_tf) {
  float sigma = 0.0f;
  if (c.mode == RPP_GLOBAL_SOFT_NMS || c.mode == RPP_PER_CLASS_SOFT_NMS) sigma = c.soft_nms_sigma;
  if (c.mode == RPP_GLOBAL_SOFT_NMS || c.mode == RPP_GLOBAL_HARD_NMS)
    *iou_thr = (sigma == 0.0f) ? 1.0f : c.iou_threshold;   // `1.0 if not sigma else iou` (B1, B2)
  else
    *iou_thr = (sigma != 0.0f) ? 1.0f : c.iou_threshold;   // `1.0 if sigma else iou` (B3)
  *sigma_tf = sigma / 2.0f;                                 // soft_nms_sigma = sigma / 2 (B4)
}

// CombinedNMS / PerClass*: per-(image, class) problems over the columns of x [B,n,C], then the per-image merge.
int per_class_chunk(Handle* h, Arena& ar, const float* x, int is_logit, const float4* deltas, const float4* boxes,
                    int q, int B, long n, long k_lim, int row0_mode, int tie_is_rank, const Outputs& out,
                    cudaStream_t st, cudaStream_t st2, cudaEvent_t ev, const Levels* levels = nullptr) {
  const rpp_config& c = h->cfg;
  const int C = c.num_classes, M = c.max_detections;
  ProblemSet ps{};
  ps.x = x; ps.is_logit = is_logit; ps.B = B; ps.n = n; ps.C = C;
  ps.deltas = deltas; ps.boxes = boxes; ps.q = q;
  ps.levels = levels;
  ps.k_lim = k_lim; ps.M = M;
  ps.score_threshold = c.score_threshold;
  ps.T_min = is_logit ? h->T_logit : std::nextafter(c.score_threshold, INFINITY);
  ps.tie_is_rank = tie_is_rank;
  const bool tpu = tpu_branch(c);
  if (tpu) {   // _tpu_per_class_hard_nms (:288-379)
    ps.consumer = RPP_CONSUME_PADDED;
    ps.padded = 2;
    ps.row0_mode = row0_mode;
    ps.clip_before = 1;
    ps.iou_threshold = c.iou_threshold;
    ps.M_lim = M;
  } else if (c.mode == RPP_COMBINED_NMS) {
    ps.consumer = RPP_CONSUME_HARD;
    ps.clip_before = 0;
    ps.iou_threshold = c.iou_threshold;
    ps.M_lim = (int)std::min<long>(M, k_lim);
  } else {
    float iou_thr, sigma_tf;
    nms_v5_args(c, &iou_thr, &sigma_tf);
    ps.consumer = sigma_tf > 0.0f ? RPP_CONSUME_SOFT : RPP_CONSUME_HARD;
    ps.clip_before = 1;
    ps.iou_threshold = iou_thr;
    ps.soft_sigma_tf = sigma_tf;
    ps.M_lim = M;
  }
  // cross-class bound: the merge keeps only the M best of C * M_lim boxes, so a class rarely needs more than a few
  {
    const int m1 = std::min(ps.M_lim, (M + C - 1) / C + h->probe_extra);
    ps.two_pass_m1 = (h->two_pass && C > 1 && 2 * m1 < ps.M_lim && (long)C * m1 <= 8192 && !tpu) ? m1 : 0;
  }
  int rc = run_problem_set(h, ar, ps, st, st2, ev);
  if (rc || ar.dry) return rc;
  if (ev) st = st2;
  if (tpu) {
    MergePaddedParams mq{};
    mq.C = C; mq.M = M; mq.score_threshold = c.score_threshold;
    mq.sel_key = ps.sel_key; mq.sel_box = ps.sel_box; mq.sel_cnt = ps.sel_cnt;
    mq.pad_score = ps.pad_score; mq.pad_box = ps.pad_box;
    mq.out_boxes = out.boxes; mq.out_scores = out.scores; mq.out_classes = (int*)out.classes;
    mq.out_valid = out.valid;
    launch_k(h->pdl, merge_padded_kernel, dim3(B), dim3(RPP_MERGE_NT), sizeof(MergeShared), st, mq);
    LAUNCHED();
    stage_mark(h, "merge", st);
    return RPP_OK;
  }

  MergeParams mp{};
  mp.C = C; mp.M = M; mp.combined = c.mode == RPP_COMBINED_NMS;
  mp.sel_key = ps.sel_key; mp.sel_box = ps.sel_box; mp.sel_cnt = ps.sel_cnt;
  memset(&mp.lv, 0, sizeof(mp.lv));
  if (levels) mp.lv = *levels;
  else { mp.lv.L = 1; mp.lv.off[1] = n; mp.lv.x[0] = x; mp.lv.d[0] = deltas; }
  mp.is_logit = is_logit; mp.N = n;
  mp.anchors = h->d_anchors; mp.boxes = boxes; mp.q = q; mp.dp = h->dp;
  mp.row0_mode = row0_mode;
  mp.score_nonneg = c.score_threshold >= 0.0f;
  mp.out_boxes = out.boxes; mp.out_scores = out.scores; mp.out_classes = out.classes; mp.out_valid = out.valid;
  size_t merge_smem = sizeof(MergeShared) + (size_t)C * M * sizeof(u64);
  mp.keys_in_smem = merge_smem <= 200 * 1024;
  if (!mp.keys_in_smem) merge_smem = sizeof(MergeShared);
  launch_k(h->pdl, merge_kernel, dim3(B), dim3(RPP_MERGE_NT), merge_smem, st, mp);
  LAUNCHED();
  stage_mark(h, "merge", st);
  return RPP_OK;
}

// Splits the batch into up to 4 image chunks: the HBM-bound collect of chunk i+1 (caller's stream) runs over the
// issue-bound NMS + merge of chunk i (the handle's side stream).  Fork/join with events only: the caller's stream
// is ordered after everything, there is no host synchronisation.
int per_class_pipeline(Handle* h, Arena& ar, const float* x, int is_logit, const float4* deltas, const float4* boxes,
                       int q, int B, long n, long k_lim, int row0_mode, int tie_is_rank, const Outputs& out,
                       cudaStream_t st, const Levels* levels = nullptr) {
  const int C = h->cfg.num_classes, M = h->cfg.max_detections;
  int nchunks = 1;
  if (h->overlap && !h->timing && B >= 16 && !levels) nchunks = B >= 32 ? 4 : 2;
  if (nchunks == 1) return per_class_chunk(h, ar, x, is_logit, deltas, boxes, q, B, n, k_lim, row0_mode, tie_is_rank,
                                           out, st, nullptr, nullptr, levels);
  int b0 = 0;
  for (int i = 0; i < nchunks; ++i) {
    const int bc = (B - b0) / (nchunks - i);
    Outputs o{};
    if (!ar.dry) {
      o.boxes = out.boxes + (size_t)b0 * M;
      o.scores = out.scores + (size_t)b0 * M;
      o.classes = (char*)out.classes + (size_t)b0 * M * 4;   // f32 / i32 classes in the per-class modes
      o.valid = out.valid + b0;
    }
    int rc = per_class_chunk(h, ar, x ? x + (size_t)b0 * n * C : nullptr, is_logit,
                             deltas ? deltas + (size_t)b0 * n : nullptr,
                             boxes ? boxes + (size_t)b0 * n * q : nullptr, q, bc, n, k_lim, row0_mode, tie_is_rank, o,
                             st, h->side, h->ev_chunk[i]);
    if (rc) return rc;
    b0 += bc;
  }
  if (!ar.dry) {
    CUDA_OK(cudaEventRecord(h->ev_join, h->side));
    CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
  }
  return RPP_OK;
}

// Rows of the Global* NMS input that were resolved without gathering them (rpp_global.cuh): row maxima (scores) and
// decoded boxes per filtered row, and the sorted keys that map a filtered row back to its anchor.
struct GlobalPre {
  const float* mraw;      // [B][k] max-class score per row
  const u64* row_keys;    // [B][k] emitted keys (tie = anchor * C + class)
  const Levels* src;      // the logits (class lookup of the selected rows)
  long N;
};

// Global*: NonMaxSuppressionV5 per image on the row maxima of x [B,n,C].
// score_rowmax: x holds logits but the row maxima are scored right here (sigmoid is monotone: max score = score of the
// max logit) and the NMS runs on that dense score column; only the class lookup of the <= M selected rows goes back
// to the logits.
int global_pipeline(Handle* h, Arena& ar, const float* x, int is_logit, const float4* deltas, const float4* boxes,
                    int B, long n, const Outputs& out, cudaStream_t st, bool score_rowmax = false,
                    const GlobalPre* pre = nullptr, const Levels* levels = nullptr) {
  const rpp_config& c = h->cfg;
  const int C = c.num_classes, M = c.max_detections;
  float* mraw = pre ? const_cast<float*>(pre->mraw) : ar.take<float>((size_t)B * n);
  if (!ar.dry && !pre) {
    const size_t rows = (size_t)B * n;
    size_t grid = (rows * 32 + 255) / 256;
    if (grid > (size_t)h->sm_count * 16) grid = (size_t)h->sm_count * 16;
    if (levels) {   // per-level pieces and / or 16-bit elements, read in place
      rowmax_levels_kernel<<<(unsigned)grid, 256, 0, st>>>(*levels, n, C, rows, mraw);
    } else if (C <= 16) {
      size_t g2 = std::min<size_t>((rows + 255) / 256, (size_t)h->sm_count * 16);
      rowmax_small_kernel<<<(unsigned)g2, 256, 0, st>>>(x, rows, C, mraw);
    } else {
      rowmax_kernel<<<(unsigned)grid, 256, 0, st>>>(x, rows, C, mraw);
    }
    LAUNCHED();
    if (score_rowmax) {
      sigmoid_kernel<<<(unsigned)std::min<size_t>((rows + 255) / 256, (size_t)h->sm_count * 16), 256, 0, st>>>(mraw, mraw,
                                                                                                            rows);
      LAUNCHED();
    }
    stage_mark(h, "rows:rowmax", st);
  }
  const int x_is_logit = pre ? 1 : is_logit;
  if (score_rowmax || pre) is_logit = 0;   // the problems below see a dense score column
  float iou_thr, sigma_tf;
  nms_v5_args(c, &iou_thr, &sigma_tf);
  ProblemSet ps{};
  ps.x = mraw; ps.is_logit = is_logit; ps.B = B; ps.n = n; ps.C = 1;
  ps.deltas = deltas; ps.boxes = boxes; ps.q = 1;
  if (pre) { ps.row_keys = pre->row_keys; ps.k_rows = n; ps.C_src = C; ps.delta_lv = pre->src; }
  else if (levels) ps.delta_lv = levels;   // x is the derived [B,n] column; the deltas stay in their pieces
  ps.k_lim = n; ps.M = M; ps.M_lim = M;
  ps.score_threshold = c.score_threshold;
  ps.T_min = is_logit ? h->T_logit : std::nextafter(c.score_threshold, INFINITY);
  ps.tie_is_rank = 0;
  ps.consumer = sigma_tf > 0.0f ? RPP_CONSUME_SOFT : RPP_CONSUME_HARD;
  ps.clip_before = 1;
  ps.iou_threshold = iou_thr;
  ps.soft_sigma_tf = sigma_tf;
  const bool tpu = tpu_branch(c);
  if (tpu) {   // _tpu_global_hard_nms (:381-432): the real IoU threshold, padded-NMS arithmetic
    ps.consumer = RPP_CONSUME_PADDED;
    ps.padded = 1;
    ps.iou_threshold = c.iou_threshold;
  }
  int rc = run_problem_set(h, ar, ps, st);
  if (rc || ar.dry) return rc;
  GlobalOutParams gp{};
  gp.tpu = tpu ? 1 : 0;
  gp.M = M; gp.sel_key = ps.sel_key; gp.sel_box = ps.sel_box; gp.sel_cnt = ps.sel_cnt;
  memset(&gp.lv, 0, sizeof(gp.lv));
  if (pre) {
    gp.lv = *pre->src;
    gp.row_keys = pre->row_keys; gp.k_rows = n;
  } else if (levels) {
    gp.lv = *levels;
  } else {
    gp.lv.L = 1; gp.lv.off[1] = n; gp.lv.x[0] = x; gp.lv.d[0] = deltas;
  }
  gp.is_logit = x_is_logit; gp.n = n; gp.C = C;
  gp.deltas = deltas; gp.anchors = h->d_anchors; gp.boxes = boxes; gp.dp = h->dp;
  gp.out_boxes = out.boxes; gp.out_scores = out.scores; gp.out_classes = (long long*)out.classes;
  gp.out_valid = out.valid;
  launch_k(h->pdl, global_out_kernel, dim3(B), dim3(128), 0, st, gp);
  LAUNCHED();
  stage_mark(h, "merge:global_out", st);
  return RPP_OK;
}

// Sorted top-k keys of every column of x [B,n,C] (C = 1 with n = rows*classes for the global filter).
int topk_keys(Handle* h, Arena& ar, const float* x, int is_logit, int B, long n, int C, long k, u64** emit_key,
              cudaStream_t st, const Levels* levels = nullptr, const GlobalTopDirectParams* direct = nullptr,
              int** emit_done = nullptr) {
  ProblemSet ps{};
  ps.direct = direct;
  ps.x = x; ps.is_logit = is_logit; ps.B = B; ps.n = n; ps.C = C;
  ps.levels = levels;
  ps.q = 1;
  ps.consumer = RPP_CONSUME_EMIT;
  ps.k_lim = k; ps.M = 1; ps.M_lim = 1;
  ps.score_threshold = -INFINITY;   // tf.nn.top_k has no threshold
  ps.T_min = -INFINITY;
  int rc = run_problem_set(h, ar, ps, st);
  *emit_key = ps.emit_key;
  if (emit_done) *emit_done = ps.emit_done;
  return rc;
}

// GenerateDetections on dense scores [B,n,C] / boxes [B,n,q,4]
int nms_dense(Handle* h, Arena& ar, const float* scores, const float4* boxes, int B, long n, int q, const Outputs& out,
              cudaStream_t st) {
  if (is_per_class_mode(h->cfg.mode))
    return per_class_pipeline(h, ar, scores, 0, nullptr, boxes, q, B, n, n, 0, 0, out, st);
  return global_pipeline(h, ar, scores, 0, nullptr, boxes, B, n, out, st);
}

// FilterTopKDetections on dense scores / boxes
// (K_out / j_off / idx_off: see the gather kernels; K_out = 0 means "the call's own k")
int topk_dense(Handle* h, Arena& ar, const float* scores, const float4* boxes, int B, long n, float* scores_out,
               float4* boxes_out, int* idx_out, cudaStream_t st, long K_out = 0, long j_off = 0, int idx_off = 0) {
  const rpp_config& c = h->cfg;
  const int C = c.num_classes;
  u64* keys = nullptr;
  if (c.filter_per_class) {
    const long k = std::min<long>(c.pre_nms_top_k, n);
    if (K_out <= 0) K_out = k;
    int rc = topk_keys(h, ar, scores, 0, B, n, C, k, &keys, st);
    if (rc || ar.dry) return rc;
    const size_t tot = (size_t)B * k * C;
    size_t grid = std::min<size_t>((tot + 255) / 256, (size_t)h->sm_count * 16);
    topk_gather_per_class_kernel<<<(unsigned)grid, 256, 0, st>>>(keys, boxes, B, n, C, k, scores_out, boxes_out,
                                                                idx_out, K_out, j_off, idx_off);
    LAUNCHED();
    return RPP_OK;
  }
  if ((double)n * C >= 2147483647.0) return fail(RPP_EINVAL, "rows x classes must stay below 2^31 for the global filter");
  const long k = std::min<long>(c.pre_nms_top_k, n * C);
  if (K_out <= 0) K_out = k;
  int rc = topk_keys(h, ar, scores, 0, B, n * C, 1, k, &keys, st);
  if (rc || ar.dry) return rc;
  const size_t tot = (size_t)B * k * C;
  size_t grid = std::min<size_t>((tot + 255) / 256, (size_t)h->sm_count * 16);
  topk_gather_global_kernel<<<(unsigned)grid, 256, 0, st>>>(keys, scores, boxes, B, n, C, k, scores_out, boxes_out,
                                                           idx_out, K_out, j_off, idx_off);
  LAUNCHED();
  return RPP_OK;
}

// add_post_processing_stage fused: logits + deltas -> detections.  The inputs are either the fused fp32 tensors
// (logits / deltas) or `levels`: the per-level pieces and / or 16-bit elements, read where they lie.
int detect_pipeline(Handle* h, Arena& ar, const float4* deltas, const float* logits, int B, const Outputs& out,
                    cudaStream_t st, const Levels* levels = nullptr) {
  const rpp_config& c = h->cfg;
  const int C = c.num_classes;
  const long N = h->N;
  const bool filtered = c.pre_nms_top_k > 0;
  const bool per_class = is_per_class_mode(c.mode);
  if (!per_class && filtered && c.filter_per_class)
    return fail(RPP_ECOMBO, "Global* NMS modes need inference.filter_per_class=false (per-class filtered boxes are "
                            "4-D; the reference fails with a rank error)");
  if (levels && per_class && (!filtered || c.filter_per_class) &&
      (C % (levels->dtype == RPP_DT_F32 ? 4 : 8) != 0 || C > 392))
    return fail(RPP_EINVAL, "rpp_detect_levels / rpp_detect_typed with the per-class filter or no filter need "
                            "num_classes % 4 == 0 (% 8 for 16-bit inputs) and at most 392 classes; fuse / convert and "
                            "call rpp_detect otherwise");
  if (!filtered) {   // TransformBoxesAndScores -> GenerateDetections on all N rows
    if (per_class) return per_class_pipeline(h, ar, logits, 1, deltas, nullptr, 1, B, N, N, 0, 0, out, st, levels);
    return global_pipeline(h, ar, logits, 1, deltas, nullptr, B, N, out, st, false, nullptr, levels);
  }
  if (c.filter_per_class) {
    const long k = std::min<long>(c.pre_nms_top_k, N);
    return per_class_pipeline(h, ar, logits, 1, deltas, nullptr, 1, B, N, k, 1, 1, out, st, levels);
  }
  // global filter (:149-161): top-k over the flat (anchor, class) axis on raw logits ...
  if ((double)N * C >= 2147483647.0) return fail(RPP_EINVAL, "anchors x classes must stay below 2^31 for the global filter");
  const long k = std::min<long>(c.pre_nms_top_k, N * C);
  // the source as Levels (class lookups, box deltas) and as its flat view (one column of N * C elements per image)
  Levels lv, flat;
  memset(&lv, 0, sizeof(lv));
  if (levels) lv = *levels;
  else { lv.L = 1; lv.off[1] = N; lv.x[0] = logits; lv.d[0] = deltas; }
  flat = lv;
  for (int l = 0; l <= lv.L; ++l) flat.off[l] = lv.off[l] * C;
  for (int l = 0; l < RPP_MAX_LEVELS; ++l) flat.d[l] = nullptr;
  u64* keys = nullptr;
  float iou_thr = 0.0f, sigma_tf = 0.0f;
  if (!per_class) nms_v5_args(c, &iou_thr, &sigma_tf);
  const int M = c.max_detections;
  const bool top_only = !per_class && sigma_tf == 0.0f && !tpu_branch(c) && M <= 1024;   // GlobalHardNMS: no suppression (B1)
  // GlobalHardNMS: one block per image turns the candidate list into the detections (global_top_direct_kernel); the
  // emission / rows / top kernels below only see the images it could not serve
  const bool direct = top_only && h->top_direct;
  GlobalTopDirectParams gd{};
  int* skip = nullptr;
  if (direct) {
    gd.src = lv; gd.C = C; gd.N = N; gd.M = M; gd.anchors = h->d_anchors; gd.dp = h->dp;
    gd.score_threshold = c.score_threshold;
    gd.debug = getenv("RPP_GTD_DEBUG") ? 1 : 0;
    gd.out_boxes = out.boxes; gd.out_scores = out.scores; gd.out_classes = (long long*)out.classes;
    gd.out_valid = out.valid;
  }
  int rc = topk_keys(h, ar, levels ? nullptr : logits, 1, B, N * C, 1, k, &keys, st, levels ? &flat : nullptr,
                     direct ? &gd : nullptr, direct ? &skip : nullptr);
  if (rc) return rc;
  if (!per_class) {
    // ... Global*: the rows the reference gathers are never materialised (rpp_global.cuh): row maxima, boxes and the
    // NonMaxSuppressionV5 order come straight from the sorted keys
    const bool soft = sigma_tf > 0.0f && !tpu_branch(c) && k <= RPP_GS_MAXK;
    u32* first = ar.take<u32>((size_t)B * N);
    float* mraw = ar.take<float>((size_t)B * k);
    float4* box_spill = soft ? ar.take<float4>((size_t)B * k) : nullptr;
    u64* skey = (soft || top_only) ? ar.take<u64>((size_t)B * k) : nullptr;
    u64* dkey = (soft || top_only) ? ar.take<u64>((size_t)B * k) : nullptr;
    int* sd_cnt = (soft || top_only) ? ar.take<int>((size_t)B * 2) : nullptr;
    if (!ar.dry) {
      GlobalRowsParams rp{};
      rp.skip = skip; rp.init_first = 1;   // every block presets its image's `first` itself: no memset node in the chain
      rp.n_loop = direct ? B : 0;
      const unsigned rows_grid = direct ? (unsigned)std::min(B, h->sm_count) : (unsigned)B;
      rp.emit_key = keys; rp.k = k; rp.C = C; rp.N = N;
      rp.first = first; rp.mraw = mraw; rp.skey = skey; rp.dkey = dkey; rp.sd_cnt = sd_cnt;
      rp.score_threshold = c.score_threshold;
      launch_k(h->pdl, global_rows_kernel, dim3(rows_grid), dim3(RPP_GROWS_NT), 0, st, rp);
      LAUNCHED();
      stage_mark(h, "rows:resolve", st);
    }
    if (soft) {
      GlobalSoftParams gs{};
      gs.k = k;
      gs.ring_cap = 2;
      while (gs.ring_cap < k) gs.ring_cap <<= 1;
      gs.M = M;
      // boxes of the best rows live in shared memory, as many as fit next to the ring (the rest is read through L2)
      const size_t budget = 200 * 1024;
      const size_t fixed = global_soft_smem(k, gs.ring_cap, 0, M);
      if (fixed > budget) return fail(RPP_EINVAL, "max_detections too large for the global soft-NMS kernel");
      gs.box_cap = (int)std::min<size_t>((size_t)k, (budget - fixed) / 16);
      if (ar.dry) return RPP_OK;
      gs.box_spill = box_spill; gs.anchors = h->d_anchors; gs.dp = h->dp;
      gs.skey = skey; gs.dkey = dkey; gs.sd_cnt = sd_cnt;
      gs.score_threshold = c.score_threshold;
      gs.soft_scale = -0.5f / sigma_tf;
      gs.iou_threshold = iou_thr;
      gs.soft_ignores_iou = c.soft_ignores_iou_threshold;
      gs.emit_key = keys; gs.lv = lv; gs.C = C; gs.N = N;
      gs.debug = getenv("RPP_GS_DEBUG") ? 1 : 0;
      gs.out_boxes = out.boxes; gs.out_scores = out.scores; gs.out_classes = (long long*)out.classes;
      gs.out_valid = out.valid;
      launch_k(h->pdl, global_soft_kernel, dim3(B), dim3(RPP_GS_NT), global_soft_smem(k, gs.ring_cap, gs.box_cap, M), st, gs);
      LAUNCHED();
      stage_mark(h, "nms:global_soft", st);
      return RPP_OK;
    }
    if (top_only) {
      if (ar.dry) return RPP_OK;
      GlobalTopParams tp{};
      tp.k = k; tp.M = M; tp.skey = skey; tp.dkey = dkey; tp.sd_cnt = sd_cnt;
      tp.emit_key = keys; tp.lv = lv; tp.C = C; tp.N = N; tp.anchors = h->d_anchors; tp.dp = h->dp;
      tp.out_boxes = out.boxes; tp.out_scores = out.scores; tp.out_classes = (long long*)out.classes;
      tp.out_valid = out.valid;
      tp.skip = skip;
      tp.n_loop = direct ? B : 0;
      launch_k(h->pdl, global_top_kernel, dim3(direct ? (unsigned)std::min(B, h->sm_count) : (unsigned)B), dim3(RPP_GTOP_NT), 0, st, tp);
      LAUNCHED();
      stage_mark(h, "nms:global_top", st);
      return RPP_OK;
    }
    GlobalPre pre{mraw, keys, &lv, N};
    return global_pipeline(h, ar, nullptr, 0, nullptr, nullptr, B, k, out, st, false, &pre);
  }
  // ... per-class modes behind the global filter: the k selected rows are transformed (sigmoid of the whole row,
  // decoded box) and fed to GenerateDetections as in the reference
  float* row_scores = ar.take<float>((size_t)B * k * C);
  float4* row_boxes = ar.take<float4>((size_t)B * k);
  if (!ar.dry) {
    int tpr = 1;                          // threads per gathered row (as in the kernel; narrow rows: per element)
    while (tpr < C && tpr < 32) tpr <<= 1;
    if (C < 16) tpr = C;
    size_t grid = std::min<size_t>(((size_t)B * k * tpr + 255) / 256, (size_t)h->sm_count * 16);
    fused_global_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(keys, lv, h->d_anchors, h->dp, B, N, C, k,
                                                            /*apply_sigmoid=*/1, row_scores, row_boxes);
    LAUNCHED();
    stage_mark(h, "rows:gather", st);
  }
  return nms_dense(h, ar, row_scores, row_boxes, B, k, 1, out, st);
}

// The handle's anchor table lives on the device that was current at rpp_create: calls must come from that device.
int check_device(const Handle* h) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) return fail(RPP_ECUDA, "cudaGetDevice failed");
  if (cur != h->device)
    return fail(RPP_EINVAL, "handle belongs to CUDA device %d but device %d is current", h->device, cur);
  return RPP_OK;
}

template <class F>
int with_arena(void* ws, size_t ws_bytes, F body) {
  Arena dry{nullptr, 0, true};
  int rc = body(dry);
  if (rc) return rc;
  if (dry.off > ws_bytes) return fail(RPP_EWORKSPACE, "workspace: need %zu bytes, got %zu", dry.off, ws_bytes);
  if (!ws && dry.off) return fail(RPP_EINVAL, "null workspace");
  Arena real{(char*)ws, 0, false};
  return body(real);
}

}  // namespace

namespace {
void host_path_free(Handle* h);
}

extern "C" {

const char* rpp_last_error(void) { return g_err; }
int rpp_last_launch_count(void) { return g_launches; }

int rpp_create(const rpp_config* cfg, void** handle) {
  if (!cfg || !handle) return fail(RPP_EINVAL, "null argument");
  *handle = nullptr;
  if (cfg->mode < 0 || cfg->mode > 4)
    return fail(RPP_EMODE, "Requested unsupported mode: %d, available modes are: CombinedNMS, GlobalSoftNMS, "
                           "GlobalHardNMS, PerClassSoftNMS, PerClassHardNMS", cfg->mode);
  const int levels = cfg->max_level - cfg->min_level + 1;
  if (cfg->H <= 0 || cfg->W <= 0 || levels <= 0 || levels > 16 || cfg->min_level < 0 || cfg->max_level > 30)
    return fail(RPP_EINVAL, "bad input_shape / levels");
  if (cfg->num_classes <= 0) return fail(RPP_EINVAL, "num_classes must be positive");
  if (cfg->n_areas < levels || cfg->n_areas > 16 || !cfg->areas)
    return fail(RPP_EINVAL, "anchor_params.areas needs one entry per level (got %d for %d levels)", cfg->n_areas,
                levels);
  if (cfg->n_ratios <= 0 || cfg->n_ratios > 8 || cfg->n_scales <= 0 || cfg->n_scales > 8 || !cfg->aspect_ratios ||
      !cfg->scales)
    return fail(RPP_EINVAL, "anchor_params.aspect_ratios / scales: 1..8 entries each");
  if (cfg->max_detections <= 0 || cfg->max_detections > 1024)
    return fail(RPP_EINVAL, "max_detections must be in 1..1024");
  if ((cfg->mode == RPP_GLOBAL_SOFT_NMS || cfg->mode == RPP_PER_CLASS_SOFT_NMS) && !(cfg->soft_nms_sigma == cfg->soft_nms_sigma))
    return fail(RPP_EINVAL, "soft_nms_sigma is required for the soft NMS modes");
  if (cfg->tpu_semantics != 0 && cfg->tpu_semantics != 1) return fail(RPP_EINVAL, "tpu_semantics must be 0 or 1");
  if (cfg->tpu_semantics && cfg->mode != RPP_GLOBAL_HARD_NMS && cfg->mode != RPP_PER_CLASS_HARD_NMS)
    return fail(RPP_EMODE, "Requested mode not supported on Cloud TPUs. Please use `GlobalHardNMS` or "
                           "`PerClassHardNMS`");   // postprocessing_ops.py:202-206
  for (int i = 0; i < 6; ++i)
    if (cfg->reserved[i] != 0) return fail(RPP_EINVAL, "rpp_config.reserved must be zero");
  if (tpu_branch(*cfg) && !(cfg->iou_threshold > 0.0f))
    return fail(RPP_EINVAL, "tpu_semantics needs iou_threshold > 0 (non_max_suppression_padded suppresses at "
                            "iou >= threshold: a non-positive threshold suppresses everything)");

  int dev = 0, count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(RPP_ECUDA, "no CUDA device: libretinapost has no CPU path");
  CUDA_OK(cudaGetDevice(&dev));
  Handle* h = new (std::nothrow) Handle();
  if (!h) return fail(RPP_EINVAL, "out of host memory");
  h->d_anchors = nullptr;
  h->d_scan_count = nullptr;
  h->side = nullptr;
  h->ev_join = nullptr;
  for (int i = 0; i < 8; ++i) h->ev_chunk[i] = nullptr;
  const int rc = [&]() -> int {
  h->cfg = *cfg;
  h->areas.assign(cfg->areas, cfg->areas + cfg->n_areas);
  h->ratios.assign(cfg->aspect_ratios, cfg->aspect_ratios + cfg->n_ratios);
  h->scales.assign(cfg->scales, cfg->scales + cfg->n_scales);
  h->cfg.areas = h->areas.data();
  h->cfg.aspect_ratios = h->ratios.data();
  h->cfg.scales = h->scales.data();
  h->levels = levels;
  h->device = dev;
  h->force_scan = 0;
  {
    const char* v = getenv("RPP_COLLECT_VARIANT");
    h->collect_variant = v ? atoi(v) : 0;
    if (h->collect_variant < 0 || h->collect_variant > 2) h->collect_variant = 0;
  }
  h->timing = 0;
  h->timed_calls = 0;
  h->ev_used = 0;
  {
    const char* v = getenv("RPP_OVERLAP");
    h->overlap = v ? atoi(v) : 0;   // measured slower on B200 (NMS blocks starve beside the persistent collect CTAs)
    v = getenv("RPP_TILE_ROWS");
    h->tile_rows = v ? atol(v) : 0;
    v = getenv("RPP_EMIT_SHORT");
    h->emit_short = v ? atoi(v) : 0;
    v = getenv("RPP_HALF_VARIANT");
    h->half_variant = v ? atoi(v) : 1;   // measured: 0.162 ms vs 0.168 ms for the 0.79 GB bf16 stream of configs[1]
    v = getenv("RPP_WARP_PROBE");
    h->warp_probe = v ? atoi(v) : 1;
    v = getenv("RPP_PROBE_EXTRA");
    h->probe_extra = v ? atoi(v) : 3;
    if (h->probe_extra < 1) h->probe_extra = 1;
    v = getenv("RPP_TWO_PASS");
    h->two_pass = v ? atoi(v) : 1;
    v = getenv("RPP_PDL");
    h->pdl = v ? atoi(v) : 1;
    v = getenv("RPP_FINISH_ARGMAX");
    h->finish_argmax = v ? atoi(v) : 1;
    v = getenv("RPP_TOP_DIRECT");
    h->top_direct = v ? atoi(v) : 1;
    v = getenv("RPP_COLLECT_CTAS");
    h->collect_ctas = v ? atoi(v) : 0;
    h->overlap_hint = 0;
    v = getenv("RPP_TARGET");
    h->target = v ? atoi(v) : kTarget;
    if (h->target < 64) h->target = 64;
    if (h->target > 3072) h->target = 3072;
  }
  CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) CUDA_OK(cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming));
  CUDA_OK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));

  AnchorParams& ap = h->ap;
  memset(&ap, 0, sizeof(ap));
  ap.H = cfg->H; ap.W = cfg->W; ap.min_level = cfg->min_level; ap.num_levels = levels;
  ap.n_ratios = cfg->n_ratios; ap.n_scales = cfg->n_scales;
  const int A = cfg->n_ratios * cfg->n_scales;
  long n = 0;
  ap.bounds[0] = 0;
  for (int li = 0; li < levels; ++li) {
    const double s = std::pow(2.0, cfg->min_level + li);
    const int fh = (int)std::ceil(cfg->H / s), fw = (int)std::ceil(cfg->W / s);  // anchor_generator.py:42-49
    ap.fw[li] = fw;
    n += (long)fh * fw * A;
    ap.bounds[li + 1] = n;
    ap.areas[li] = cfg->areas[li];
  }
  for (int i = 0; i < cfg->n_ratios; ++i) ap.ratios[i] = cfg->aspect_ratios[i];
  for (int i = 0; i < cfg->n_scales; ++i) ap.scales[i] = cfg->scales[i];
  h->N = n;
  if (n <= 0 || n >= 0x7fffffffL) return fail(RPP_EINVAL, "anchor count out of range");

  h->dp.shape[0] = (float)cfg->H; h->dp.shape[1] = (float)cfg->W;
  h->dp.shape[2] = (float)cfg->H; h->dp.shape[3] = (float)cfg->W;
  for (int i = 0; i < 4; ++i) h->dp.var[i] = cfg->box_variance[i];
  h->dp.scale = cfg->scale_box_targets;

  cudaError_t e = cudaMalloc(&h->d_anchors, (size_t)n * sizeof(float4));
  if (e != cudaSuccess) return fail(RPP_ECUDA, "cudaMalloc anchors: %s", cudaGetErrorString(e));
  anchors_kernel<<<(unsigned)((n + 255) / 256), 256>>>(ap, n, h->d_anchors);
  if (cudaMalloc(&h->d_scan_count, sizeof(u32)) == cudaSuccess) cudaMemset(h->d_scan_count, 0, sizeof(u32));
  float* d_t = nullptr;
  e = cudaMalloc(&d_t, sizeof(float));
  if (e == cudaSuccess) {
    logit_threshold_kernel<<<1, 32>>>(cfg->score_threshold, d_t);
    e = cudaMemcpy(&h->T_logit, d_t, sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_t);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RPP_ECUDA, "create kernels: %s", cudaGetErrorString(e));
  cudaFuncSetAttribute(col_problem_kernel<RPP_CONSUME_HARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(col_problem_kernel<RPP_CONSUME_SOFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(col_problem_kernel<RPP_CONSUME_EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(col_problem_kernel<RPP_CONSUME_PADDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(merge_padded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(global_soft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(emit_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(global_top_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(sample_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols4_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols8_half_kernel<4, RPP_DT_F16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols8_half_kernel<4, RPP_DT_BF16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols8_half_kernel<8, RPP_DT_F16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols8_half_kernel<8, RPP_DT_BF16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols4_levels_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols4_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_colsv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(collect_cols4_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return RPP_OK;
  }();
  if (rc != RPP_OK) {   // nothing leaks on a failed create
    rpp_destroy(h);
    return rc;
  }
  *handle = h;
  return RPP_OK;
}

int rpp_destroy(void* handle) {
  Handle* h = (Handle*)handle;
  if (!h) return RPP_OK;
  cudaFree(h->d_anchors);
  cudaFree(h->d_scan_count);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  host_path_free(h);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (int i = 0; i < 8; ++i) if (h->ev_chunk[i]) cudaEventDestroy(h->ev_chunk[i]);
  delete h;
  return RPP_OK;
}

long rpp_num_anchors(void* handle) { return handle ? ((Handle*)handle)->N : -1; }
int rpp_num_levels(void* handle) { return handle ? ((Handle*)handle)->levels : -1; }
int rpp_anchor_boundaries(void* handle, long* h_out) {
  Handle* h = (Handle*)handle;
  if (!h || !h_out) return fail(RPP_EINVAL, "null argument");
  for (int i = 0; i <= h->levels; ++i) h_out[i] = h->ap.bounds[i];
  return RPP_OK;
}
int rpp_classes_itemsize(void* handle) {
  Handle* h = (Handle*)handle;
  if (!h) return -1;
  if (tpu_branch(h->cfg)) return 4;   // both TPU branches cast the classes to int32 (:375, :425-426)
  return (h->cfg.mode == RPP_GLOBAL_SOFT_NMS || h->cfg.mode == RPP_GLOBAL_HARD_NMS) ? 8 : 4;
}
int rpp_debug_force_exact_scan(void* handle, int on) {
  if (!handle) return fail(RPP_EINVAL, "null handle");
  ((Handle*)handle)->force_scan = on ? 1 : 0;
  return RPP_OK;
}

int rpp_debug_exact_scans(void* handle, unsigned long long* h_count, int reset) {
  Handle* h = (Handle*)handle;
  if (!h || !h_count) return fail(RPP_EINVAL, "null argument");
  if (int rc = check_device(h)) return rc;
  u32 v = 0;
  if (h->d_scan_count) {
    CUDA_OK(cudaMemcpy(&v, h->d_scan_count, sizeof(u32), cudaMemcpyDeviceToHost));   // (synchronises the device)
    if (reset) CUDA_OK(cudaMemset(h->d_scan_count, 0, sizeof(u32)));
  }
  *h_count = v;
  return RPP_OK;
}

int rpp_debug_sample_plan(long n, int C, long k_lim, int emit, int* h_out) {
  if (!h_out || n <= 0 || C <= 0) return fail(RPP_EINVAL, "bad argument");
  bool fine = false;
  int target = 0;
  const SamplePlan p = choose_plan(n, C, emit != 0, k_lim, kTarget, 0, &fine, &target);
  h_out[0] = p.on ? 1 : 0; h_out[1] = p.stride; h_out[2] = p.G; h_out[3] = p.rows_per_group; h_out[4] = p.rank;
  h_out[5] = p.CAP; h_out[6] = target; h_out[7] = fine ? 1 : 0;
  return RPP_OK;
}

int rpp_debug_stage_timing(void* handle, int on) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(RPP_EINVAL, "null handle");
  h->timing = on ? 1 : 0;
  h->timed_calls = 0;
  h->ev_used = 0;
  return RPP_OK;
}

namespace {
// Sums the recorded segments by label (first-seen order), averaged per API call; resets the recording.
int stage_collect(Handle* h, std::vector<std::pair<const char*, double>>& out, int* n_calls) {
  out.clear();
  const int n = h->timed_calls;
  if (n_calls) *n_calls = n;
  if (n == 0 || h->ev_used == 0) { h->ev_used = 0; h->timed_calls = 0; return RPP_OK; }
  CUDA_OK(cudaEventSynchronize(h->events[h->ev_used - 1]));
  for (size_t i = 1; i < h->ev_used; ++i) {
    const char* lab = h->ev_label[i];
    if (!lab) continue;
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, h->events[i - 1], h->events[i]));
    size_t j = 0;
    while (j < out.size() && strcmp(out[j].first, lab) != 0) ++j;
    if (j == out.size()) out.push_back({lab, 0.0});
    out[j].second += ms / n;
  }
  h->ev_used = 0;
  h->timed_calls = 0;
  return RPP_OK;
}
}  // namespace

int rpp_debug_stage_ms(void* handle, float* h_ms, int* n_calls) {
  Handle* h = (Handle*)handle;
  if (!h || !h_ms) return fail(RPP_EINVAL, "null argument");
  for (int s = 0; s < kStages; ++s) h_ms[s] = 0.f;
  std::vector<std::pair<const char*, double>> segs;
  if (int rc = stage_collect(h, segs, n_calls)) return rc;
  for (auto& sg : segs) {   // the four classic buckets; everything else (emission, row gathers) counts as "nms"
    const char* l = sg.first;
    int bkt = 2;
    if (!strncmp(l, "sample", 6)) bkt = 0;
    else if (!strncmp(l, "collect", 7)) bkt = 1;
    else if (!strncmp(l, "merge", 5)) bkt = 3;
    h_ms[bkt] += (float)sg.second;
  }
  return RPP_OK;
}

int rpp_debug_stage_report(void* handle, char* buf, int cap, int* n_calls) {
  Handle* h = (Handle*)handle;
  if (!h || !buf || cap <= 0) return fail(RPP_EINVAL, "null argument");
  std::vector<std::pair<const char*, double>> segs;
  if (int rc = stage_collect(h, segs, n_calls)) return rc;
  int off = 0;
  buf[0] = 0;
  for (auto& sg : segs) {
    const int w = snprintf(buf + off, (size_t)(cap - off), "%s%s=%.6f", off ? ";" : "", sg.first, sg.second);
    if (w < 0 || w >= cap - off) return fail(RPP_EINVAL, "stage report buffer too small");
    off += w;
  }
  return RPP_OK;
}

int rpp_anchors(void* handle, float* d_out, void* stream) {
  Handle* h = (Handle*)handle;
  if (!h || !d_out) return fail(RPP_EINVAL, "null argument");
  CUDA_OK(cudaMemcpyAsync(d_out, h->d_anchors, (size_t)h->N * sizeof(float4), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  return RPP_OK;
}

size_t rpp_workspace_bytes(void* handle, int B, long n) {
  Handle* h = (Handle*)handle;
  if (!h || B <= 0) return 0;
  const Outputs none{};
  size_t need = 0;
  // the handle's role is not known here (rpp_detect, rpp_nms or rpp_topk): size for the largest of those that are
  // valid for this config
  {
    Arena a{nullptr, 0, true};
    if (n <= 0 && detect_pipeline(h, a, nullptr, nullptr, B, none, nullptr) == RPP_OK) need = std::max(need, a.off);
  }
  if (n <= 0 && (double)h->N * h->cfg.num_classes < 2147483647.0) {   // rpp_efficient_nms
    Arena a{nullptr, 0, true};
    u64* keys = nullptr;
    if (topk_keys(h, a, nullptr, 1, B, h->N * h->cfg.num_classes, 1,
                  std::min<long>(RPP_EFFNMS_SELECTED, h->N * h->cfg.num_classes), &keys, nullptr) == RPP_OK)
      need = std::max(need, a.off);
  }
  if (n > 0) {
    const bool per_class = is_per_class_mode(h->cfg.mode);
    for (int q = 1; q <= (per_class ? 2 : 1); ++q) {
      Arena a{nullptr, 0, true};
      if (nms_dense(h, a, nullptr, nullptr, B, n, q == 1 ? 1 : h->cfg.num_classes, none, nullptr) == RPP_OK)
        need = std::max(need, a.off);
    }
    if (h->cfg.pre_nms_top_k > 0) {
      Arena a{nullptr, 0, true};
      if (topk_dense(h, a, nullptr, nullptr, B, n, nullptr, nullptr, nullptr, nullptr) == RPP_OK)
        need = std::max(need, a.off);
    }
  }
  return need + 4096;
}

int rpp_decode(void* handle, const float* d_logits, const float* d_deltas, int B, float* d_scores, float* d_boxes,
               void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || B <= 0) return fail(RPP_EINVAL, "bad argument");
  if (int rc = check_device(h)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h->cfg.num_classes;
  if (d_scores) {
    if (!d_logits) return fail(RPP_EINVAL, "null logits");
    const size_t n = (size_t)B * h->N * C;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)h->sm_count * 16) grid = (size_t)h->sm_count * 16;
    sigmoid_kernel<<<(unsigned)grid, 256, 0, st>>>(d_logits, d_scores, n);
    LAUNCHED();
  }
  if (d_boxes) {
    if (!d_deltas) return fail(RPP_EINVAL, "null deltas");
    const size_t n = (size_t)B * h->N;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)h->sm_count * 16) grid = (size_t)h->sm_count * 16;
    decode_kernel<<<(unsigned)grid, 256, 0, st>>>((const float4*)d_deltas, h->d_anchors, B, h->N, h->dp,
                                                  (float4*)d_boxes);
    LAUNCHED();
  }
  return RPP_OK;
}

int rpp_topk(void* handle, const float* d_scores, const float* d_boxes, int B, long n, float* d_scores_out,
             float* d_boxes_out, int* d_index_out, void* ws, size_t ws_bytes, void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !d_scores || !d_boxes || !d_scores_out || !d_boxes_out || B <= 0 || n <= 0 || n >= 0x7fffffffL)
    return fail(RPP_EINVAL, "bad argument");
  if (h->cfg.pre_nms_top_k <= 0) return fail(RPP_EINVAL, "pre_nms_top_k must be positive for rpp_topk");
  if (int rc = check_device(h)) return rc;
  if (h->timing) ++h->timed_calls;
  return with_arena(ws, ws_bytes, [&](Arena& ar) {
    return topk_dense(h, ar, d_scores, (const float4*)d_boxes, B, n, d_scores_out, (float4*)d_boxes_out, d_index_out,
                      (cudaStream_t)stream);
  });
}

int rpp_topk_levels(void* handle, int n_levels, const float* const* d_scores_levels,
                    const float* const* d_boxes_levels, const long* n_rows, int B, float* d_scores_out,
                    float* d_boxes_out, int* d_index_out, void* ws, size_t ws_bytes, void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !d_scores_levels || !d_boxes_levels || !n_rows || !d_scores_out || !d_boxes_out || B <= 0 ||
      n_levels <= 0 || n_levels > RPP_MAX_LEVELS)
    return fail(RPP_EINVAL, "bad argument");
  if (h->cfg.pre_nms_top_k <= 0) return fail(RPP_EINVAL, "pre_nms_top_k must be positive for rpp_topk_levels");
  if (int rc = check_device(h)) return rc;
  const rpp_config& c = h->cfg;
  long K = 0, n_tot = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (!d_scores_levels[l] || !d_boxes_levels[l] || n_rows[l] <= 0 || n_rows[l] >= 0x7fffffffL)
      return fail(RPP_EINVAL, "bad level %d", l);
    K += std::min<long>(c.pre_nms_top_k, c.filter_per_class ? n_rows[l] : n_rows[l] * c.num_classes);
    n_tot += n_rows[l];
  }
  if ((double)n_tot * c.num_classes >= 2147483647.0) return fail(RPP_EINVAL, "rows x classes must stay below 2^31");
  if (h->timing) ++h->timed_calls;
  int launches = 0;
  long j_off = 0, idx_off = 0;
  for (int l = 0; l < n_levels; ++l) {   // one filter per segment, stream-ordered over the same scratch
    const long n = n_rows[l];
    int rc = with_arena(ws, ws_bytes, [&](Arena& ar) {
      return topk_dense(h, ar, d_scores_levels[l], (const float4*)d_boxes_levels[l], B, n, d_scores_out,
                        (float4*)d_boxes_out, d_index_out, (cudaStream_t)stream, K, j_off, (int)idx_off);
    });
    if (rc) return rc;
    launches += g_launches;
    g_launches = 0;
    j_off += std::min<long>(c.pre_nms_top_k, c.filter_per_class ? n : n * c.num_classes);
    idx_off += n;
  }
  g_launches = launches;
  return RPP_OK;
}

int rpp_nms(void* handle, const float* d_scores, const float* d_boxes, int B, long n, int q, float* d_boxes_out,
            float* d_scores_out, void* d_classes_out, int* d_valid_out, void* ws, size_t ws_bytes, void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !d_scores || !d_boxes || B <= 0 || n <= 0 || n >= 0x7fffffffL) return fail(RPP_EINVAL, "bad argument");
  if (int rc = check_device(h)) return rc;
  if (h->timing) ++h->timed_calls;
  const rpp_config& c = h->cfg;
  if (q != 1 && q != c.num_classes) return fail(RPP_EINVAL, "boxes must be [B,n,4] or [B,n,num_classes,4]");
  if (!is_per_class_mode(c.mode) && q != 1)
    return fail(RPP_ECOMBO, "Global* NMS modes need class-agnostic [B,n,4] boxes (inference.filter_per_class=false)");
  const Outputs out{(float4*)d_boxes_out, d_scores_out, d_classes_out, d_valid_out};
  return with_arena(ws, ws_bytes, [&](Arena& ar) {
    return nms_dense(h, ar, d_scores, (const float4*)d_boxes, B, n, q, out, (cudaStream_t)stream);
  });
}

static int detect_impl(Handle* h, const float* d_deltas, const float* d_logits, int B, float* d_boxes_out,
                       float* d_scores_out, void* d_classes_out, int* d_valid_out, void* ws, size_t ws_bytes,
                       void* stream) {
  if (!h || !d_deltas || !d_logits || B <= 0) return fail(RPP_EINVAL, "bad argument");
  if (int rc = check_device(h)) return rc;
  if (h->timing) ++h->timed_calls;
  const Outputs out{(float4*)d_boxes_out, d_scores_out, d_classes_out, d_valid_out};
  return with_arena(ws, ws_bytes, [&](Arena& ar) {
    return detect_pipeline(h, ar, (const float4*)d_deltas, d_logits, B, out, (cudaStream_t)stream);
  });
}

int rpp_coco_format(void* handle, const float* d_boxes, const float* d_scores, const void* d_classes,
                    const int* d_valid, int B, const float* d_resize_scale, const int* d_class_map, int* d_bbox_out,
                    int* d_category_out, float* d_score_out, int* d_image_out, int* d_total_out, void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !d_boxes || !d_scores || !d_classes || !d_valid || !d_bbox_out || !d_category_out || !d_score_out ||
      !d_image_out || !d_total_out || B <= 0)
    return fail(RPP_EINVAL, "bad argument");
  CocoParams cp{};
  cp.boxes = (const float4*)d_boxes; cp.scores = d_scores; cp.classes = d_classes; cp.valid = d_valid;
  cp.class_kind = h->cfg.mode == RPP_COMBINED_NMS ? 0 : (rpp_classes_itemsize(h) == 8 ? 1 : 2);
  cp.B = B; cp.M = h->cfg.max_detections;
  cp.resize_scale = d_resize_scale;
  cp.in_h = (float)h->cfg.H; cp.in_w = (float)h->cfg.W;
  cp.class_map = d_class_map; cp.num_classes = h->cfg.num_classes;
  cp.bbox_out = (int4*)d_bbox_out; cp.category_out = d_category_out; cp.score_out = d_score_out;
  cp.image_out = d_image_out; cp.total_out = d_total_out;
  coco_format_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(cp);
  LAUNCHED();
  return RPP_OK;
}

int rpp_efficient_nms(void* handle, const float* d_raw_boxes, const float* d_class_logits, const float* d_anchor_boxes,
                      int B, int* d_valid_detections, float* d_detection_boxes, float* d_detection_scores,
                      int* d_detection_classes, void* ws, size_t ws_bytes, void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !d_raw_boxes || !d_class_logits || !d_valid_detections || !d_detection_boxes || !d_detection_scores ||
      !d_detection_classes || B <= 0)
    return fail(RPP_EINVAL, "bad argument");
  if (int rc = check_device(h)) return rc;
  const int C = h->cfg.num_classes, M = h->cfg.max_detections;
  const long N = h->N;
  if ((double)N * C >= 2147483647.0) return fail(RPP_EINVAL, "anchors x classes out of range");
  if (h->timing) ++h->timed_calls;
  cudaStream_t st = (cudaStream_t)stream;
  return with_arena(ws, ws_bytes, [&](Arena& ar) {
    const long k = std::min<long>(RPP_EFFNMS_SELECTED, N * C);
    u64* keys = nullptr;
    int rc = topk_keys(h, ar, d_class_logits, 1, B, N * C, 1, k, &keys, st);
    if (rc || ar.dry) return rc;
    EffNmsParams ep{};
    ep.emit_key = keys; ep.k = k;
    ep.deltas = (const float4*)d_raw_boxes;
    ep.anchors = d_anchor_boxes ? (const float4*)d_anchor_boxes : h->d_anchors;
    ep.N = N; ep.C = C;
    ep.score_threshold = h->cfg.score_threshold; ep.iou_threshold = h->cfg.iou_threshold;
    ep.M = M;
    ep.out_valid = d_valid_detections; ep.out_boxes = (float4*)d_detection_boxes;
    ep.out_scores = d_detection_scores; ep.out_classes = d_detection_classes;
    effnms_kernel<<<B, RPP_NMS_NT, (size_t)M * 24, st>>>(ep);
    LAUNCHED();
    stage_mark(h, "nms:effnms", st);
    return (int)RPP_OK;
  });
}

int rpp_detect_typed(void* handle, int n_pieces, const void* const* d_deltas, const void* const* d_logits, int dtype,
                     int B, float* d_boxes_out, float* d_scores_out, void* d_classes_out, int* d_valid_out, void* ws,
                     size_t ws_bytes, void* stream) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !d_deltas || !d_logits || B <= 0) return fail(RPP_EINVAL, "bad argument");
  if (dtype != RPP_DT_F32 && dtype != RPP_DT_F16 && dtype != RPP_DT_BF16) return fail(RPP_EINVAL, "bad dtype");
  if (int rc = check_device(h)) return rc;
  if (n_pieces != 1 && n_pieces != h->levels)
    return fail(RPP_EINVAL, "n_pieces must be 1 (fused tensors) or the number of levels (%d)", h->levels);
  if (n_pieces > RPP_MAX_LEVELS) return fail(RPP_EINVAL, "too many levels");
  if (h->timing) ++h->timed_calls;
  Levels lv;
  memset(&lv, 0, sizeof(lv));
  lv.L = n_pieces;
  lv.dtype = dtype;
  for (int l = 0; l < n_pieces; ++l) {
    if (!d_deltas[l] || !d_logits[l]) return fail(RPP_EINVAL, "null tensor pointer");
    if (((uintptr_t)d_deltas[l] % 16) || ((uintptr_t)d_logits[l] % 16))
      return fail(RPP_EINVAL, "tensors must be 16-byte aligned");
    lv.off[l] = n_pieces == 1 ? 0 : h->ap.bounds[l];
    lv.x[l] = (const float*)d_logits[l];
    lv.d[l] = (const float4*)d_deltas[l];
  }
  lv.off[n_pieces] = h->N;
  const Outputs out{(float4*)d_boxes_out, d_scores_out, d_classes_out, d_valid_out};
  return with_arena(ws, ws_bytes, [&](Arena& ar) {
    return detect_pipeline(h, ar, nullptr, nullptr, B, out, (cudaStream_t)stream, &lv);
  });
}

int rpp_detect_levels(void* handle, const float* const* d_deltas_levels, const float* const* d_logits_levels, int B,
                      float* d_boxes_out, float* d_scores_out, void* d_classes_out, int* d_valid_out, void* ws,
                      size_t ws_bytes, void* stream) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(RPP_EINVAL, "bad argument");
  return rpp_detect_typed(handle, h->levels, (const void* const*)d_deltas_levels,
                          (const void* const*)d_logits_levels, RPP_DT_F32, B, d_boxes_out, d_scores_out,
                          d_classes_out, d_valid_out, ws, ws_bytes, stream);
}

int rpp_detect(void* handle, const float* d_deltas, const float* d_logits, int B, float* d_boxes_out,
               float* d_scores_out, void* d_classes_out, int* d_valid_out, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  return detect_impl((Handle*)handle, d_deltas, d_logits, B, d_boxes_out, d_scores_out, d_classes_out, d_valid_out,
                     ws, ws_bytes, stream);
}

}  // extern "C"

namespace {
void host_path_free(Handle* h) {
  Handle::HostPath& hp = h->hp;
  for (int i = 0; i < 2; ++i) {
    cudaFree(hp.d_logits[i]); cudaFree(hp.d_deltas[i]);
    if (hp.ready[i]) cudaEventDestroy(hp.ready[i]);
    if (hp.done[i]) cudaEventDestroy(hp.done[i]);
  }
  cudaFree(hp.ws); cudaFree(hp.d_boxes); cudaFree(hp.d_scores); cudaFree(hp.d_classes); cudaFree(hp.d_valid);
  if (hp.s_copy) cudaStreamDestroy(hp.s_copy);
  if (hp.s_comp) cudaStreamDestroy(hp.s_comp);
  hp = Handle::HostPath();
}
}  // namespace

extern "C" {

int rpp_detect_host_typed(void* handle, int device, const void* h_deltas, const void* h_logits, int dtype, int B,
                          float* h_boxes, float* h_scores, void* h_classes, int* h_valid) {
  Handle* h = (Handle*)handle;
  g_launches = 0;
  if (!h || !h_deltas || !h_logits || !h_boxes || !h_scores || !h_classes || !h_valid || B <= 0)
    return fail(RPP_EINVAL, "bad argument");
  if (dtype != RPP_DT_F32 && dtype != RPP_DT_F16 && dtype != RPP_DT_BF16) return fail(RPP_EINVAL, "bad dtype");
  const size_t esz = dtype == RPP_DT_F32 ? 4 : 2;
  if (device != h->device)
    return fail(RPP_EINVAL, "handle belongs to CUDA device %d, not device %d", h->device, device);
  CUDA_OK(cudaSetDevice(device));
  Handle::HostPath& hp = h->hp;
  const int C = h->cfg.num_classes, M = h->cfg.max_detections;
  const size_t lg_img = (size_t)h->N * C * esz, dl_img = (size_t)h->N * 4 * esz;
  // staging is sized for fp32 so that a handle can serve both element types
  const size_t lg_img32 = (size_t)h->N * C * sizeof(float), dl_img32 = (size_t)h->N * 4 * sizeof(float);
  // chunk: ~256 MB of logits per staging buffer, at least one image
  int chunk = (int)std::max<size_t>(1, (256u << 20) / lg_img32);
  if (chunk > B) chunk = B;
  const int csz = rpp_classes_itemsize(h);
  if (hp.chunk < chunk || hp.B_out < B) {
    host_path_free(h);
    CUDA_OK(cudaStreamCreateWithFlags(&hp.s_copy, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&hp.s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_OK(cudaEventCreateWithFlags(&hp.ready[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&hp.done[i], cudaEventDisableTiming));
      CUDA_OK(cudaMalloc(&hp.d_logits[i], lg_img32 * chunk));
      CUDA_OK(cudaMalloc(&hp.d_deltas[i], dl_img32 * chunk));
    }
    hp.ws_bytes = rpp_workspace_bytes(h, chunk, 0);
    CUDA_OK(cudaMalloc(&hp.ws, hp.ws_bytes));
    CUDA_OK(cudaMalloc(&hp.d_boxes, (size_t)B * M * 4 * sizeof(float)));
    CUDA_OK(cudaMalloc(&hp.d_scores, (size_t)B * M * sizeof(float)));
    CUDA_OK(cudaMalloc(&hp.d_classes, (size_t)B * M * csz));
    CUDA_OK(cudaMalloc(&hp.d_valid, (size_t)B * sizeof(int)));
    hp.chunk = chunk;
    hp.B_out = B;
  }
  chunk = hp.chunk < B ? hp.chunk : B;
  int i = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++i) {
    const int bc = B - b0 < chunk ? B - b0 : chunk;
    const int buf = i & 1;
    if (i >= 2) CUDA_OK(cudaStreamWaitEvent(hp.s_copy, hp.done[buf], 0));
    CUDA_OK(cudaMemcpyAsync(hp.d_logits[buf], (const char*)h_logits + lg_img * b0, lg_img * bc,
                            cudaMemcpyHostToDevice, hp.s_copy));
    CUDA_OK(cudaMemcpyAsync(hp.d_deltas[buf], (const char*)h_deltas + dl_img * b0, dl_img * bc,
                            cudaMemcpyHostToDevice, hp.s_copy));
    CUDA_OK(cudaEventRecord(hp.ready[buf], hp.s_copy));
    CUDA_OK(cudaStreamWaitEvent(hp.s_comp, hp.ready[buf], 0));
    int rc;
    if (dtype == RPP_DT_F32) {
      rc = detect_impl(h, hp.d_deltas[buf], hp.d_logits[buf], bc, hp.d_boxes + (size_t)b0 * M * 4,
                       hp.d_scores + (size_t)b0 * M, (char*)hp.d_classes + (size_t)b0 * M * csz, hp.d_valid + b0,
                       hp.ws, hp.ws_bytes, hp.s_comp);
    } else {
      const void* dp[1] = {hp.d_deltas[buf]};
      const void* lp[1] = {hp.d_logits[buf]};
      const int launched = g_launches;
      rc = rpp_detect_typed(h, 1, dp, lp, dtype, bc, hp.d_boxes + (size_t)b0 * M * 4, hp.d_scores + (size_t)b0 * M,
                            (char*)hp.d_classes + (size_t)b0 * M * csz, hp.d_valid + b0, hp.ws, hp.ws_bytes,
                            hp.s_comp);
      g_launches += launched;
    }
    if (rc) return rc;
    CUDA_OK(cudaEventRecord(hp.done[buf], hp.s_comp));
  }
  CUDA_OK(cudaMemcpyAsync(h_boxes, hp.d_boxes, (size_t)B * M * 4 * sizeof(float), cudaMemcpyDeviceToHost, hp.s_comp));
  CUDA_OK(cudaMemcpyAsync(h_scores, hp.d_scores, (size_t)B * M * sizeof(float), cudaMemcpyDeviceToHost, hp.s_comp));
  CUDA_OK(cudaMemcpyAsync(h_classes, hp.d_classes, (size_t)B * M * csz, cudaMemcpyDeviceToHost, hp.s_comp));
  CUDA_OK(cudaMemcpyAsync(h_valid, hp.d_valid, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, hp.s_comp));
  CUDA_OK(cudaStreamSynchronize(hp.s_comp));
  return RPP_OK;
}

int rpp_detect_host(void* handle, int device, const float* h_deltas, const float* h_logits, int B, float* h_boxes,
                    float* h_scores, void* h_classes, int* h_valid) {
  return rpp_detect_host_typed(handle, device, h_deltas, h_logits, RPP_DT_F32, B, h_boxes, h_scores, h_classes,
                               h_valid);
}

}  // extern "C"
