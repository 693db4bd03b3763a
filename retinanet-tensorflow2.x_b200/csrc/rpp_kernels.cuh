// rpp_kernels.cuh — device kernels of libretinapost (sm_100a).  See DESIGN.md for the data layout and the
// per-kernel roofline; reference citations are relative to /root/reference/retinanet/.
#pragma once
#include "rpp_common.cuh"
#include "rpp_select.cuh"

// ===============================================================================================================
// The [B, N, C] logit tensor and the [B, N, 4] delta tensor, either fused (L = 1) or as the per-level head outputs
// the model produces ([B, H_l, W_l, A*C] NHWC = [B, n_l, C]; FuseDetections, postprocessing_ops.py:15-56, only
// concatenates them).  Reading the levels in place saves the 1.65 GB concat copy per batch.
// ===============================================================================================================
#define RPP_MAX_LEVELS 8
#define RPP_DT_F32 0
#define RPP_DT_F16 1
#define RPP_DT_BF16 2
struct Levels {
  int L;
  int dtype;                           // element type of x and d: the reference casts whatever the heads emit to
                                       // fp32 first (postprocessing_ops.py:111-112); f16 / bf16 are converted on
                                       // load (exactly), so half-precision heads stream half the bytes
  long off[RPP_MAX_LEVELS + 1];        // cumulative rows: level l holds rows [off[l], off[l+1])
  const float* x[RPP_MAX_LEVELS];      // [B, n_l, C]
  const float4* d[RPP_MAX_LEVELS];     // [B, n_l] float4 (may be null for score tensors)
  int tile_off[RPP_MAX_LEVELS + 1];    // collect kernel: cumulative tiles per image
};
__device__ __forceinline__ int lv_find(const Levels& lv, long r) {
  int l = 0;
  while (l + 1 < lv.L && r >= lv.off[l + 1]) ++l;
  return l;
}
__device__ __forceinline__ float half_bits_to_f32(unsigned short h, int dtype) {
  return dtype == RPP_DT_F16 ? __half2float(__ushort_as_half(h)) : __uint_as_float((u32)h << 16);
}
// element (b, r, c) as fp32, any dtype (for half types x[] really points at 16-bit data)
__device__ __forceinline__ float lv_val(const Levels& lv, int b, long r, int C, int c) {
  const int l = lv_find(lv, r);
  const size_t idx = ((size_t)b * (lv.off[l + 1] - lv.off[l]) + (r - lv.off[l])) * C + c;
  if (lv.dtype == RPP_DT_F32) return __ldg(lv.x[l] + idx);
  return half_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(lv.x[l]) + idx), lv.dtype);
}
__device__ __forceinline__ float4 lv_delta(const Levels& lv, int b, long r) {
  const int l = lv_find(lv, r);
  const size_t idx = (size_t)b * (lv.off[l + 1] - lv.off[l]) + (r - lv.off[l]);
  if (lv.dtype == RPP_DT_F32) return lv.d[l][idx];
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(lv.d[l]) + idx);   // 4 x 16 bit
  return make_float4(half_bits_to_f32((unsigned short)(h.x & 0xffffu), lv.dtype),
                     half_bits_to_f32((unsigned short)(h.x >> 16), lv.dtype),
                     half_bits_to_f32((unsigned short)(h.y & 0xffffu), lv.dtype),
                     half_bits_to_f32((unsigned short)(h.y >> 16), lv.dtype));
}

// ===============================================================================================================
// K0a  anchors — AnchorBoxGenerator (dataloader/anchor_generator.py:24-104).  One thread per anchor.
// ===============================================================================================================
struct AnchorParams {
  int H, W, min_level, num_levels, n_ratios, n_scales;
  long bounds[17];     // cumulative anchors per level
  int fw[16];          // feature width per level
  double areas[16];
  double ratios[8];
  double scales[8];
};

__global__ void anchors_kernel(AnchorParams ap, long N, float4* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int li = 0;
  while (li + 1 < ap.num_levels && i >= ap.bounds[li + 1]) ++li;
  const int A = ap.n_ratios * ap.n_scales;
  const long r = i - ap.bounds[li];
  const int a = (int)(r % A);
  const long cell = r / A;
  const int x = (int)(cell % ap.fw[li]), y = (int)(cell / ap.fw[li]);
  const float stride = (float)(1 << (ap.min_level + li));
  // _compute_dims :51-63 — area/ratio in float64 (Python), then fp32 sqrt / div / mul; ratio-major, scale-minor
  const int ri = a / ap.n_scales, si = a % ap.n_scales;
  const float h = __fsqrt_rn((float)(ap.areas[li] / ap.ratios[ri]));
  const float w = __fdiv_rn((float)ap.areas[li], h);
  const float s = (float)ap.scales[si];
  out[i] = make_float4(__fmul_rn(__fadd_rn((float)x, 0.5f), stride), __fmul_rn(__fadd_rn((float)y, 0.5f), stride),
                       __fmul_rn(s, w), __fmul_rn(s, h));
}

// ===============================================================================================================
// K0b  exact pre-image of the score threshold: the smallest logit x with sigmoid(x) > score_threshold, found by
// bisection over the ordered float encoding with the device sigmoid itself, so `logit >= T` <=> `score > thr`.
// ===============================================================================================================
__global__ void logit_threshold_kernel(float score_threshold, float* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // ordered encodings of -inf .. +inf
  u32 lo = ord_f32(-INFINITY), hi = ord_f32(INFINITY);
  if (sigmoid_f32(unord_f32(lo)) > score_threshold) { *out = -INFINITY; return; }
  if (!(sigmoid_f32(unord_f32(hi)) > score_threshold)) { *out = INFINITY; return; }  // callers treat +inf as "none"
  // invariant: S(lo) <= thr < S(hi)
  while (hi - lo > 1u) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (sigmoid_f32(unord_f32(mid)) > score_threshold) hi = mid; else lo = mid;
  }
  *out = unord_f32(hi);
}

// ===============================================================================================================
// K0c  TransformBoxesAndScores.call materialised (postprocessing_ops.py:107-117) — stage-wise entry rpp_decode.
// ===============================================================================================================
__global__ void sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = sigmoid_f32(x[i]);
}

__global__ void decode_kernel(const float4* __restrict__ deltas, const float4* __restrict__ anchors, long B, long N,
                              DecodeParams dp, float4* __restrict__ out) {
  const size_t tot = (size_t)B * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
    out[i] = decode_box(deltas[i], anchors[i % N], dp);
}

// ===============================================================================================================
// K1  sample -> per-(image, class) pre-threshold.
//
// Candidates for one NMS problem are "the best few hundred of a column of N logits".  A strided sample of the
// column (every `stride`-th anchor, dealt round-robin into G groups) gives G group maxima; their r-th smallest is
// an estimate of the logit whose upper tail holds ~target elements.  The estimate only has to be roughly right:
// the problem kernel consumes candidates lazily and falls back to an exact scan of the column if the list runs
// dry, so results never depend on it.
//   K1a  sample_max_kernel   grid (B, SPLIT): thread = (class, row lane) keeps RPP_GPT group maxima in registers
//                            over its share of the rounds (row-contiguous loads, RPP_GPT independent loads in
//                            flight), then merges them into gm[b][g][c] with atomicMax.
//   K1b  sample_rank_kernel  grid B: r-th smallest of the G maxima per class -> T[b*C + c] = max(est, T_min).
// ===============================================================================================================
#define RPP_GPT 8   // groups per thread; G = lanes * RPP_GPT

template <bool LEVELS, bool HALF>
__global__ void __launch_bounds__(1024, 2)   // two 960-thread blocks per SM: at most 32 registers
sample_max_kernel(Levels lv /*[B,N,C]*/, long N, int C, int stride, int lanes,
                                  int rounds, u32* __restrict__ gm /*[B][G][C]*/) {
  const int b = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
  const int c = threadIdx.x % C, rl = threadIdx.x / C;
  // LEVELS: the table is indexed at run time, so it is staged in shared memory (run-time indexing of kernel
  // parameters costs a select chain per access); the sampled rows of a thread only grow -> running level cursor
  __shared__ long s_off[RPP_MAX_LEVELS + 1];
  __shared__ const float* s_x[RPP_MAX_LEVELS];
  if (LEVELS) {
    if (threadIdx.x <= lv.L) s_off[threadIdx.x] = lv.off[threadIdx.x];
    if (threadIdx.x < lv.L) s_x[threadIdx.x] = lv.x[threadIdx.x];
    __syncthreads();
  }
  if (rl >= lanes) return;
  const int G = lanes * RPP_GPT;
  float m[RPP_GPT];
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i) m[i] = -INFINITY;
  const float* base = lv.x[0] + (size_t)b * N * C + c;   // fused tensor (LEVELS == false)
  const unsigned short* hbase = reinterpret_cast<const unsigned short*>(lv.x[0]) + (size_t)b * N * C + c;
  const int dtype = lv.dtype;
  const int nlv = lv.L;
  int lvl = 0;
  for (int r = split; r < rounds; r += nsplit) {
    float v[RPP_GPT];
    // LEVELS: a round covers G * stride consecutive rows; when they all lie in one level (all but the few rounds that
    // straddle a boundary) the level is resolved once and the loads look like the fused tensor's
    bool one_level = false;
    const float* lbase = nullptr;
    if (LEVELS) {
      const long row_first = (long)r * G * stride, row_last = ((long)r * G + G - 1) * stride;
      while (lvl + 1 < nlv && row_first >= s_off[lvl + 1]) ++lvl;
      one_level = row_last < s_off[lvl + 1];
      // lbase[row * C] (in elements of the input type) is element (b, row - off_l, c) of the level tensor
      const long n_l = s_off[lvl + 1] - s_off[lvl];
      lbase = s_x[lvl];
      const long shift = ((long)b * n_l - s_off[lvl]) * C + c;
      lbase = HALF ? reinterpret_cast<const float*>(reinterpret_cast<const unsigned short*>(lbase) + shift)
                   : lbase + shift;
    }
    if (LEVELS && one_level) {
#pragma unroll
      for (int i = 0; i < RPP_GPT; ++i) {
        const long s = (long)r * G + rl + i * lanes;
        if (HALF) v[i] = half_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(lbase) + (size_t)(s * stride) * C), dtype);
        else v[i] = __ldg(lbase + (size_t)(s * stride) * C);
      }
    } else
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) {
      const long s = (long)r * G + rl + i * lanes;  // sampled row index; group = rl + i * lanes
      if (LEVELS) {
        const long row = s * stride;
        while (lvl + 1 < nlv && row >= s_off[lvl + 1]) ++lvl;
        const size_t idx = ((size_t)b * (s_off[lvl + 1] - s_off[lvl]) + (row - s_off[lvl])) * C + c;
        // (the element type is a template parameter here too: a run-time branch around the load keeps the compiler
        // from batching the RPP_GPT loads of a round)
        if (HALF) v[i] = half_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(s_x[lvl]) + idx), dtype);
        else v[i] = __ldg(s_x[lvl] + idx);
      } else if (HALF) {
        v[i] = half_bits_to_f32(__ldg(hbase + (size_t)(s * stride) * C), dtype);
      } else {
        v[i] = __ldg(base + (size_t)(s * stride) * C);
      }
    }
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) m[i] = fmaxf(m[i], v[i]);
  }
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i)
    atomicMax(&gm[((size_t)b * G + rl + i * lanes) * C + c], ord_f32(m[i]));
}

// Single-column variant (C == 1: the flat anchors x classes axis of the global filter and of the EfficientNMS entry,
// n % 4 == 0): every sample is one 16-byte load = four consecutive elements, so the same number of sampled elements
// touches a quarter of the sectors (a strided sample of single floats fetches 32 bytes for every 4 it uses).
__global__ void __launch_bounds__(1024, 2)
sample_max_flat4_kernel(const float4* __restrict__ x4 /*[B][n4]*/, long n4, int stride4, int lanes, int rounds4,
                        u32* __restrict__ gm /*[B][G]*/) {
  const int b = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
  const int rl = threadIdx.x;
  if (rl >= lanes) return;
  const int G = lanes * RPP_GPT;
  float m[RPP_GPT];
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i) m[i] = -INFINITY;
  const float4* base = x4 + (size_t)b * n4;
  for (int r = split; r < rounds4; r += nsplit) {
    float4 v[RPP_GPT];
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) v[i] = __ldg(base + (size_t)((long)r * G + rl + i * lanes) * stride4);
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) m[i] = fmaxf(m[i], fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i) atomicMax(&gm[(size_t)b * G + rl + i * lanes], ord_f32(m[i]));
}

#define RPP_RANK_CPB 8   // classes per block
__global__ void sample_rank_kernel(const u32* __restrict__ gm, int C, int G, int rank, float T_min,
                                   float* __restrict__ T) {
  extern __shared__ u32 s_gm[];  // [G][RPP_RANK_CPB]
  const int b = blockIdx.x, c0 = blockIdx.y * RPP_RANK_CPB;
  const int nc = C - c0 < RPP_RANK_CPB ? C - c0 : RPP_RANK_CPB;
  for (int i = threadIdx.x; i < G * nc; i += blockDim.x) {
    const int g = i / nc, cc = i - g * nc;
    s_gm[g * RPP_RANK_CPB + cc] = gm[((size_t)b * G + g) * C + c0 + cc];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * nc; i += blockDim.x) {
    const int g0 = i / nc, cc = i - g0 * nc;
    const u32 v = s_gm[g0 * RPP_RANK_CPB + cc];
    int less = 0, eq = 0;
    for (int g = 0; g < G; ++g) {
      const u32 o = s_gm[g * RPP_RANK_CPB + cc];
      less += o < v;
      eq += o == v;
    }
    if (less <= rank && rank < less + eq) T[(size_t)b * C + c0 + cc] = fmaxf(unord_f32(v), T_min);
  }
}

// Same result with a 128-key register bitonic sort per (image, class): one warp per class (G <= 128).
__global__ void __launch_bounds__(RPP_RANK_CPB * 32)
sample_rank_sort_kernel(const u32* __restrict__ gm, int C, int G, int rank, float T_min, float* __restrict__ T) {
  __shared__ u32 s_gm[128 * RPP_RANK_CPB];
  const int b = blockIdx.x, c0 = blockIdx.y * RPP_RANK_CPB;
  const int nc = C - c0 < RPP_RANK_CPB ? C - c0 : RPP_RANK_CPB;
  for (int i = threadIdx.x; i < G * nc; i += blockDim.x) {
    const int g = i / nc, cc = i - g * nc;
    s_gm[g * RPP_RANK_CPB + cc] = gm[((size_t)b * G + g) * C + c0 + cc];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, cc = threadIdx.x >> 5;
  if (cc >= nc) return;
  u32 v[4];
#pragma unroll
  for (int sidx = 0; sidx < 4; ++sidx) {
    const int g = sidx * 32 + lane;
    v[sidx] = g < G ? s_gm[g * RPP_RANK_CPB + cc] : 0xffffffffu;   // pads sort to the end
  }
#pragma unroll
  for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int ds = j >> 5;   // slot distance 1 or 2
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx) {
          if ((sidx & ds) == 0) {
            const int e = sidx * 32 + lane;
            const bool asc = (e & size) == 0;
            const u32 a0 = v[sidx], a1 = v[sidx | ds];
            if (asc ? (a0 > a1) : (a0 < a1)) { v[sidx] = a1; v[sidx | ds] = a0; }
          }
        }
      } else {
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx) {
          const int e = sidx * 32 + lane;
          const u32 other = __shfl_xor_sync(RPP_FULL_MASK, v[sidx], j);
          const bool asc = (e & size) == 0;
          const bool low = (lane & j) == 0;
          const bool keep_min = asc == low;
          v[sidx] = keep_min ? (other < v[sidx] ? other : v[sidx]) : (other > v[sidx] ? other : v[sidx]);
        }
      }
    }
  }
  // ascending: element `rank` is the answer
  const int rs = rank >> 5, rl = rank & 31;
  u32 ans = 0u;
#pragma unroll
  for (int sidx = 0; sidx < 4; ++sidx)
    if (sidx == rs) ans = v[sidx];
  if (lane == rl) T[(size_t)b * C + c0 + cc] = fmaxf(unord_f32(ans), T_min);
}

__global__ void fill_kernel(float* p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ===============================================================================================================
// K2  collect — the HBM-bound stream.  Reads class_logits [B,N,C] exactly once with 128-bit streaming loads and
// appends every element with logit >= T[b,c] to that problem's candidate list as (logit bits, anchor index).
// No sigmoid here: the comparison is on raw logits (monotone pre-image of the score), so the kernel issues one
// LDG.128 and four compares per 16 bytes.  Thread = (class quad, row lane): its four thresholds live in registers
// for a whole tile and UNROLL independent loads are in flight per thread.  Hits (~1 %) are staged per class in
// shared memory and flushed once per tile with ONE global atomic per (tile, class); a class that overflows its
// stage appends directly.  Tiles are handed out dynamically (atomic tile counter) so the tail is balanced.
// ===============================================================================================================
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void append_cand(u32* cand_count, uint2* cand, int CAP, size_t p, float v, u32 idx) {
  const u32 slot = atomicAdd(&cand_count[p], 1u);
  if (slot < (u32)CAP) cand[p * (size_t)CAP + slot] = make_uint2(__float_as_uint(v), idx);
}

#define RPP_STAGE_CAP 64
#define RPP_COLLECT_NT 512

template <int UNROLL, int MINB>
__global__ void __launch_bounds__(RPP_COLLECT_NT, MINB)
collect_cols4_kernel(const float4* __restrict__ x4 /*[B,N,C/4]*/, const float* __restrict__ T /*[B*C]*/,
                     u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, int B, long N, int C4,
                     int lanes /*row lanes per block*/, int rows_per_tile, int tiles_per_image,
                     u32* __restrict__ tile_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = C4 * 4;
  uint2* s_stage = reinterpret_cast<uint2*>(smem_raw);   // [C][RPP_STAGE_CAP] staged (logit bits, row)
  u32* s_cnt = reinterpret_cast<u32*>(s_stage + (size_t)C * RPP_STAGE_CAP);  // [C]
  u32* s_base = s_cnt + C;                               // [C]
  __shared__ long s_tile;
  __shared__ u32 s_span;
  const int tid = threadIdx.x;
  const int cq = tid % C4, rl = tid / C4;
  const bool active = rl < lanes;
  const long n_tiles = (long)B * tiles_per_image;
  const float* x = reinterpret_cast<const float*>(x4);
  for (;;) {
    if (tid == 0) s_tile = (long)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
    if (tid == 0) s_span = 0u;
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const int b = (int)(tile / tiles_per_image);
    const long r0 = (long)(tile % tiles_per_image) * rows_per_tile;
    const long r1 = r0 + rows_per_tile < N ? r0 + rows_per_tile : N;
    const size_t pbase = (size_t)b * C;
    const float* xb = x + (size_t)b * N * C;
    if (active) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(T + (size_t)b * C) + cq);
      const float4* src = x4 + (size_t)b * N * C4 + cq;
      for (long row = r0 + rl; row < r1; row += (long)lanes * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long r = row + (long)u * lanes;
          v[u] = r < r1 ? ld_stream_f4(src + (size_t)r * C4)
                        : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        u32 mask = 0u;  // bit 4u+i: component i of load u passes its class threshold (NaN never passes >=)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          u32 m = (v[u].x >= t4.x ? 1u : 0u) | (v[u].y >= t4.y ? 2u : 0u) | (v[u].z >= t4.z ? 4u : 0u) |
                  (v[u].w >= t4.w ? 8u : 0u);
          if (row + (long)u * lanes >= r1) m = 0u;
          mask |= m << (4 * u);
        }
        // rare path (~1 % of elements).  The value is re-read by address (an L1 hit: the line was just loaded by
        // this warp) instead of being selected out of 16 registers by a run-time index.
        while (mask) {
          const int bit = __ffs(mask) - 1;
          mask &= mask - 1u;
          const int c = cq * 4 + (bit & 3);
          const u32 r = (u32)(row + (long)(bit >> 2) * lanes);
          const float val = __ldg(xb + (size_t)r * C + c);
          const u32 slot = atomicAdd(&s_cnt[c], 1u);
          if (slot < RPP_STAGE_CAP) s_stage[c * RPP_STAGE_CAP + slot] = make_uint2(__float_as_uint(val), r);
          else append_cand(cand_count, cand, CAP, pbase + c, val, r);
        }
      }
    }
    __syncthreads();
    // flush: one global atomic per class that staged anything
    for (int c = tid; c < C; c += RPP_COLLECT_NT) {
      const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
      s_base[c] = n ? atomicAdd(&cand_count[pbase + c], n) : 0u;
      if (n) atomicMax(&s_span, n);
    }
    __syncthreads();
    const int span = (int)s_span;   // the fullest class stage of this tile: copy only that many slots per class
    for (int e = tid; e < C * span; e += RPP_COLLECT_NT) {
      const int c = e / span, r = e - c * span;
      const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
      if ((u32)r < n) {
        const u32 slot = s_base[c] + (u32)r;
        if (slot < (u32)CAP) cand[(pbase + c) * (size_t)CAP + slot] = s_stage[c * RPP_STAGE_CAP + r];
      }
    }
    __syncthreads();
  }
}

// Per-level variant (rpp_detect_levels): same kernel, tiles are (image, level, row range).
template <int UNROLL, int MINB>
__global__ void __launch_bounds__(RPP_COLLECT_NT, MINB)
collect_cols4_levels_kernel(Levels lv /*[B,N,C] in per-level pieces*/, const float* __restrict__ T /*[B*C]*/,
                     u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, int B, long N, int C4,
                     int lanes /*row lanes per block*/, int rows_per_tile, int tiles_per_image,
                     u32* __restrict__ tile_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = C4 * 4;
  uint2* s_stage = reinterpret_cast<uint2*>(smem_raw);   // [C][RPP_STAGE_CAP] staged (logit bits, row)
  u32* s_cnt = reinterpret_cast<u32*>(s_stage + (size_t)C * RPP_STAGE_CAP);  // [C]
  u32* s_base = s_cnt + C;                               // [C]
  __shared__ long s_tile;
  __shared__ u32 s_span;
  __shared__ long s_nl, s_goff, s_r0;      // per-tile level geometry, resolved once by thread 0
  __shared__ const float* s_xb;
  const int tid = threadIdx.x;
  const int cq = tid % C4, rl = tid / C4;
  const bool active = rl < lanes;
  const long n_tiles = (long)B * tiles_per_image;
  for (;;) {
    if (tid == 0) {
      const long t = (long)atomicAdd(tile_counter, 1u);
      s_tile = t;
      if (t < n_tiles) {
        const int bb = (int)(t / tiles_per_image), t_img = (int)(t % tiles_per_image);
        int l = 0;
        while (l + 1 < lv.L && t_img >= lv.tile_off[l + 1]) ++l;
        s_nl = lv.off[l + 1] - lv.off[l];
        s_goff = lv.off[l];
        s_r0 = (long)(t_img - lv.tile_off[l]) * rows_per_tile;
        s_xb = lv.x[l] + (size_t)bb * (lv.off[l + 1] - lv.off[l]) * C;
      }
    }
    for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
    if (tid == 0) s_span = 0u;
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const int b = (int)(tile / tiles_per_image);
    // rows are LOCAL to the level inside the loop; `goff` turns them into fused row indices when staged
    const long n_l = s_nl, goff = s_goff, r0 = s_r0;
    const float* __restrict__ xb = s_xb;
    const long r1 = r0 + rows_per_tile < n_l ? r0 + rows_per_tile : n_l;
    const size_t pbase = (size_t)b * C;
    if (active) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(T + (size_t)b * C) + cq);
      const float4* __restrict__ src = reinterpret_cast<const float4*>(xb) + cq;
      for (long row = r0 + rl; row < r1; row += (long)lanes * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long r = row + (long)u * lanes;
          v[u] = r < r1 ? ld_stream_f4(src + (size_t)r * C4)
                        : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        u32 mask = 0u;  // bit 4u+i: component i of load u passes its class threshold (NaN never passes >=)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          u32 m = (v[u].x >= t4.x ? 1u : 0u) | (v[u].y >= t4.y ? 2u : 0u) | (v[u].z >= t4.z ? 4u : 0u) |
                  (v[u].w >= t4.w ? 8u : 0u);
          if (row + (long)u * lanes >= r1) m = 0u;
          mask |= m << (4 * u);
        }
        // rare path (~1 % of elements).  The value is re-read by address (an L1 hit: the line was just loaded by
        // this warp) instead of being selected out of 16 registers by a run-time index.
        while (mask) {
          const int bit = __ffs(mask) - 1;
          mask &= mask - 1u;
          const int c = cq * 4 + (bit & 3);
          const u32 r = (u32)(row + (long)(bit >> 2) * lanes);
          const float val = __ldg(xb + (size_t)r * C + c);
          const u32 slot = atomicAdd(&s_cnt[c], 1u);
          if (slot < RPP_STAGE_CAP) s_stage[c * RPP_STAGE_CAP + slot] = make_uint2(__float_as_uint(val), (u32)goff + r);
          else append_cand(cand_count, cand, CAP, pbase + c, val, (u32)goff + r);
        }
      }
    }
    __syncthreads();
    // flush: one global atomic per class that staged anything
    for (int c = tid; c < C; c += RPP_COLLECT_NT) {
      const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
      s_base[c] = n ? atomicAdd(&cand_count[pbase + c], n) : 0u;
      if (n) atomicMax(&s_span, n);
    }
    __syncthreads();
    const int span = (int)s_span;   // the fullest class stage of this tile: copy only that many slots per class
    for (int e = tid; e < C * span; e += RPP_COLLECT_NT) {
      const int c = e / span, r = e - c * span;
      const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
      if ((u32)r < n) {
        const u32 slot = s_base[c] + (u32)r;
        if (slot < (u32)CAP) cand[(pbase + c) * (size_t)CAP + slot] = s_stage[c * RPP_STAGE_CAP + r];
      }
    }
    __syncthreads();
  }
}

// 16-bit variant (f16 / bf16 logits, fused or per-level): one LDG.128 = 8 classes of one anchor, converted exactly to
// fp32 and compared against 8 register-resident thresholds; everything downstream sees fp32 logit bits.
// packed helpers: two 16-bit values per 32-bit word, compared natively (HSETP2) against thresholds that were rounded
// UP to the 16-bit type — for a 16-bit value v and a float T:  v >= T  <=>  v >= ceil16(T)
template <int DT> __device__ __forceinline__ u32 pack_thresholds_ru(float lo, float hi);
template <> __device__ __forceinline__ u32 pack_thresholds_ru<RPP_DT_F16>(float lo, float hi) {
  return (u32)__half_as_ushort(__float2half_ru(lo)) | ((u32)__half_as_ushort(__float2half_ru(hi)) << 16);
}
template <> __device__ __forceinline__ u32 pack_thresholds_ru<RPP_DT_BF16>(float lo, float hi) {
  return (u32)__bfloat16_as_ushort(__float2bfloat16_ru(lo)) | ((u32)__bfloat16_as_ushort(__float2bfloat16_ru(hi)) << 16);
}
template <int DT> __device__ __forceinline__ u32 ge2_mask(u32 v, u32 t);
template <> __device__ __forceinline__ u32 ge2_mask<RPP_DT_F16>(u32 v, u32 t) {
  return __hge2_mask(*reinterpret_cast<const __half2*>(&v), *reinterpret_cast<const __half2*>(&t));
}
template <> __device__ __forceinline__ u32 ge2_mask<RPP_DT_BF16>(u32 v, u32 t) {
  return __hge2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v), *reinterpret_cast<const __nv_bfloat162*>(&t));
}

template <int UNROLL, int DT, int MINB>
__global__ void __launch_bounds__(RPP_COLLECT_NT, MINB)
collect_cols8_half_kernel(Levels lv, const float* __restrict__ T /*[B*C]*/, u32* __restrict__ cand_count,
                          uint2* __restrict__ cand, int CAP, int B, long N, int C8, int lanes, int rows_per_tile,
                          int tiles_per_image, u32* __restrict__ tile_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = C8 * 8;
  uint2* s_stage = reinterpret_cast<uint2*>(smem_raw);
  u32* s_cnt = reinterpret_cast<u32*>(s_stage + (size_t)C * RPP_STAGE_CAP);
  u32* s_base = s_cnt + C;
  __shared__ long s_tile;
  __shared__ u32 s_span;
  __shared__ long s_nl, s_goff, s_r0;      // per-tile level geometry, resolved once by thread 0
  __shared__ const unsigned short* s_xb;
  const int tid = threadIdx.x;
  const int co = tid % C8, rl = tid / C8;
  const bool active = rl < lanes;
  const int dtype = lv.dtype;
  const long n_tiles = (long)B * tiles_per_image;
  for (;;) {
    if (tid == 0) {
      const long t = (long)atomicAdd(tile_counter, 1u);
      s_tile = t;
      if (t < n_tiles) {
        const int bb = (int)(t / tiles_per_image), t_img = (int)(t % tiles_per_image);
        int l = 0;
        while (l + 1 < lv.L && t_img >= lv.tile_off[l + 1]) ++l;
        s_nl = lv.off[l + 1] - lv.off[l];
        s_goff = lv.off[l];
        s_r0 = (long)(t_img - lv.tile_off[l]) * rows_per_tile;
        s_xb = reinterpret_cast<const unsigned short*>(lv.x[l]) + (size_t)bb * (lv.off[l + 1] - lv.off[l]) * C;
      }
    }
    for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
    if (tid == 0) s_span = 0u;
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const int b = (int)(tile / tiles_per_image);
    const long n_l = s_nl, goff = s_goff, r0 = s_r0;
    const long r1 = r0 + rows_per_tile < n_l ? r0 + rows_per_tile : n_l;
    const size_t pbase = (size_t)b * C;
    const unsigned short* __restrict__ xb = s_xb;
    if (active) {
      u32 th[4];   // the 8 class thresholds of this thread, packed in the input's 16-bit type
      {
        const float4 ta = __ldg(reinterpret_cast<const float4*>(T + (size_t)b * C) + 2 * co);
        const float4 tb = __ldg(reinterpret_cast<const float4*>(T + (size_t)b * C) + 2 * co + 1);
        th[0] = pack_thresholds_ru<DT>(ta.x, ta.y); th[1] = pack_thresholds_ru<DT>(ta.z, ta.w);
        th[2] = pack_thresholds_ru<DT>(tb.x, tb.y); th[3] = pack_thresholds_ru<DT>(tb.z, tb.w);
      }
      const uint4* src = reinterpret_cast<const uint4*>(xb) + co;
      for (long row = r0 + rl; row < r1; row += (long)lanes * UNROLL) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long r = row + (long)u * lanes;
          if (r < r1) {
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + (size_t)r * C8));
          } else {
            v[u] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        u64 mask = 0ull;   // bit 8u+i: class 8*co+i of load u passes
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          // 0xFFFF per passing half -> one bit per class: a byte of each half through PRMT, one distinct bit kept per
          // byte, bytes summed (= OR, the bits are distinct) by a multiply
          const u32 g0 = ge2_mask<DT>(v[u].x, th[0]), g1 = ge2_mask<DT>(v[u].y, th[1]);
          const u32 g2 = ge2_mask<DT>(v[u].z, th[2]), g3 = ge2_mask<DT>(v[u].w, th[3]);
          const u32 t = (__byte_perm(g0, g1, 0x6420) & 0x08040201u) | ((__byte_perm(g2, g3, 0x6420) & 0x08040201u) << 4);
          u32 m = (t * 0x01010101u) >> 24;   // bit 2i + h = half h of word i
          if (row + (long)u * lanes >= r1) m = 0u;
          mask |= (u64)m << (8 * u);
        }
        while (mask) {
          const int bit = __ffsll((long long)mask) - 1;
          mask &= mask - 1ull;
          const int c = co * 8 + (bit & 7);
          const long r = row + (long)(bit >> 3) * lanes;
          const float val = half_bits_to_f32(__ldg(xb + (size_t)r * C + c), dtype);
          const u32 slot = atomicAdd(&s_cnt[c], 1u);
          if (slot < RPP_STAGE_CAP) s_stage[c * RPP_STAGE_CAP + slot] = make_uint2(__float_as_uint(val), (u32)(goff + r));
          else append_cand(cand_count, cand, CAP, pbase + c, val, (u32)(goff + r));
        }
      }
    }
    __syncthreads();
    for (int c = tid; c < C; c += RPP_COLLECT_NT) {
      const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
      s_base[c] = n ? atomicAdd(&cand_count[pbase + c], n) : 0u;
      if (n) atomicMax(&s_span, n);
    }
    __syncthreads();
    const int span = (int)s_span;   // the fullest class stage of this tile: copy only that many slots per class
    for (int e = tid; e < C * span; e += RPP_COLLECT_NT) {
      const int c = e / span, r = e - c * span;
      const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
      if ((u32)r < n) {
        const u32 slot = s_base[c] + (u32)r;
        if (slot < (u32)CAP) cand[(pbase + c) * (size_t)CAP + slot] = s_stage[c * RPP_STAGE_CAP + r];
      }
    }
    __syncthreads();
  }
}

// Single-column variant (C == 1: the flat anchors x classes axis of the global filter, or the row maxima of the
// Global* modes): x [B, n], n % 4 == 0, one threshold per image.  Same structure: streaming LDG.128, hits queued in
// shared memory, one global atomic per tile.
#define RPP_FLAT_QCAP 1024
template <int UNROLL>
__global__ void __launch_bounds__(RPP_COLLECT_NT, 3)
collect_flat4_kernel(const float4* __restrict__ x4 /*[B, n/4]*/, const float* __restrict__ T /*[B]*/,
                     u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, int B, long n4,
                     int f4_per_tile, int tiles_per_image, u32* __restrict__ tile_counter) {
  __shared__ uint2 s_q[RPP_FLAT_QCAP];
  __shared__ u32 s_qn, s_base;
  __shared__ long s_tile;
  const int tid = threadIdx.x;
  const long n_tiles = (long)B * tiles_per_image;
  for (;;) {
    if (tid == 0) { s_tile = (long)atomicAdd(tile_counter, 1u); s_qn = 0u; }
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const int b = (int)(tile / tiles_per_image);
    const long f0 = (long)(tile % tiles_per_image) * f4_per_tile;
    const long f1 = f0 + f4_per_tile < n4 ? f0 + f4_per_tile : n4;
    const float t = __ldg(T + b);
    const float4* src = x4 + (size_t)b * n4;
    for (long f = f0 + tid; f < f1; f += (long)RPP_COLLECT_NT * UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long ff = f + (long)u * RPP_COLLECT_NT;
        v[u] = ff < f1 ? ld_stream_f4(src + ff) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long ff = f + (long)u * RPP_COLLECT_NT;
        if (ff >= f1) continue;
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (e[i] >= t) {
            const u32 idx = (u32)(ff * 4 + i);
            const u32 slot = atomicAdd(&s_qn, 1u);
            if (slot < RPP_FLAT_QCAP) s_q[slot] = make_uint2(__float_as_uint(e[i]), idx);
            else append_cand(cand_count, cand, CAP, (size_t)b, e[i], idx);
          }
        }
      }
    }
    __syncthreads();
    const u32 nq = s_qn < RPP_FLAT_QCAP ? s_qn : RPP_FLAT_QCAP;
    if (tid == 0) s_base = nq ? atomicAdd(&cand_count[b], nq) : 0u;
    __syncthreads();
    for (u32 i = tid; i < nq; i += RPP_COLLECT_NT) {
      const u32 slot = s_base + i;
      if (slot < (u32)CAP) cand[(size_t)b * CAP + slot] = s_q[i];
    }
    __syncthreads();
  }
}

// single column, any n / alignment: grid (chunks, B), no index arithmetic beyond the stride
__global__ void collect_flat1_kernel(const float* __restrict__ x /*[B,n]*/, const float* __restrict__ T,
                                     u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, long n) {
  const int b = blockIdx.y;
  const float t = __ldg(T + b);
  const float* xb = x + (size_t)b * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = __ldg(xb + i);
    if (v >= t) append_cand(cand_count, cand, CAP, (size_t)b, v, (u32)i);
  }
}

// generic C (C % 4 != 0, or unaligned base): one element per thread step
__global__ void collect_cols1_kernel(const float* __restrict__ x, const float* __restrict__ T,
                                     u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, int B, long N,
                                     int C) {
  const size_t tot = (size_t)B * N * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + e);
    const size_t row = e / C;
    const int c = (int)(e - row * C);
    const int b = (int)(row / N);
    const size_t p = (size_t)b * C + c;
    if (v >= T[p]) append_cand(cand_count, cand, CAP, p, v, (u32)(row - (size_t)b * N));
  }
}

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
  int* emit_done;          // [P] 1 = emit_sort_kernel already wrote this problem's keys
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

__device__ __forceinline__ float4 col_box(const ColProblemParams& P, int b, int c, u32 row) {
  if (P.boxes) {
    const int qi = P.q > 1 ? (c < P.q - 1 ? c : P.q - 1) : 0;  // boxes[:, min(q-1, c)] (:440)
    return P.boxes[((size_t)b * P.N + row) * P.q + qi];
  }
  return decode_box(lv_delta(P.lv, b, row), P.anchors[row], P.dp);
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
// Soft NMS consumer: NonMaxSuppressionV5 with soft_nms_sigma > 0, lazily re-scored exactly as the
// TF kernel does it.  The priority queue is split in two: candidates never popped yet are the not-yet-consumed part
// of the sorted stream (their order is static), and candidates popped, decayed and pushed back live in R (shared
// memory, spilling to global).  Each step pops the larger of (stream head, max of R); a popped candidate multiplies
// its score by the weights of the boxes selected since its last visit, newest first, in fp32 in exactly that order,
// stopping when it falls to the score threshold; it is selected iff the score did not change.
// expf_glibc reproduces libm's expf bit for bit (checked against glibc on 4.5e8 inputs): TF's kernel calls
// Eigen::numext::exp<float> = expf.
// ---------------------------------------------------------------------------------------------------------------
__constant__ u64 c_exp2f_tab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

__device__ __forceinline__ float expf_glibc(float x) {
  if (!(x > -87.0f && x < 88.0f)) return (float)exp((double)x);  // under/overflow tails: correctly rounded exp
  const double N = 32.0;
  const double InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
  const double z = __dmul_rn(InvLn2N, (double)x);
  double kd = __dadd_rn(z, SHIFT);
  const u64 ki = (u64)__double_as_longlong(kd);
  kd = __dsub_rn(kd, SHIFT);
  const double r = __dsub_rn(z, kd);
  const u64 t = c_exp2f_tab[ki & 31u] + (ki << 47);
  const double sc = __longlong_as_double((long long)t);
  const double zz = __dadd_rn(__dmul_rn(C0, r), C1);
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
  y = __dadd_rn(__dmul_rn(zz, r2), y);
  y = __dmul_rn(y, sc);
  return __double2float_rn(y);
}

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
      // long lists are converted in place (global memory): remember it for the finish pass
      if (keys == gkeys && tid == 0) P.cand_count[p] = n_raw | 0x80000000u;
    }
    __syncthreads();
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
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
      const float v = s_sc[i];
      if (!(v > -INFINITY)) continue;
      int rank = 0;
      for (int j = 0; j < n_all; ++j) {
        const float o = s_sc[j];
        rank += (o > v) || (o == v && j < i);
      }
      if (rank == Mtop - 1) s_L = v;   // exactly one element has this rank
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
#define RPP_PROBE_WARPS 8
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

__global__ void __launch_bounds__(RPP_PROBE_WARPS * 32) probe_warp_kernel(ColProblemParams P, size_t n_problems) {
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
      orig = col_box(P, b, c, key_tie(key));
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
__global__ void __launch_bounds__(RPP_EMIT_NT) emit_sort_kernel(ColProblemParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EmitShared* sh = reinterpret_cast<EmitShared*>(smem_raw);
  const int tid = threadIdx.x;
  const size_t p = blockIdx.x;
  const int b = (int)(p / P.C), c = (int)(p % P.C);
  if (tid == 0) { P.emit_done[p] = 0; sh->valid = 0; }
  const u32 n_raw = P.cand_count[p];
  if (n_raw & 0x80000000u) return;
  const bool list_ok = !(P.force_scan || n_raw > (u32)P.CAP || n_raw > RPP_EMIT_CAP || n_raw == 0);
  int n_keys = 0;   // scored keys in sh->keys[0 .. n_keys)
  int nv = 0;       // of which valid (strictly above everything that is not in sh->keys)
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
    if (P.force_scan) return;   // debug: the generic kernel's exact scan is what is being tested
    __syncthreads();
    u64 KBr = ~0ull;
    u32 population = 0u;
    long want = P.k_lim + P.k_lim / 8 + 64;
    if (want > RPP_EMIT_CAP) return;   // more than one block's worth: the generic kernel takes it
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
    if ((long)nv < P.k_lim) return;   // a huge tie group at the cut, or a coarse radix cut: the generic kernel decides
  }
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

struct GlobalOutParams {
  int M;
  const u64* sel_key;     // [B][M]  (score | ~row)
  const float4* sel_box;  // [B][M]
  const int* sel_cnt;     // [B]
  const float* x;         // [B,n,C] logits or scores
  int is_logit; long n; int C;
  const float4* deltas; const float4* anchors; const float4* boxes; DecodeParams dp;
  float4* out_boxes; float* out_scores; long long* out_classes; int* out_valid;
  int tpu;                // _tpu_global_hard_nms (:402-431): int32 classes, -1 in every field beyond valid
};

__global__ void global_out_kernel(GlobalOutParams P) {
  const int b = blockIdx.x;
  const int valid = P.sel_cnt[b];
  if (threadIdx.x == 0) P.out_valid[b] = valid;
  for (int i = threadIdx.x; i < P.M; i += blockDim.x) {
    const size_t o = (size_t)b * P.M + i;
    if (i < valid) {
      const u64 k = P.sel_key[o];
      const u32 row = key_tie(k);
      const float* xr = P.x + ((size_t)b * P.n + row) * P.C;
      // tf.argmax over scores: first class whose SCORE equals the row maximum
      float best = -INFINITY;
      for (int c = 0; c < P.C; ++c) best = fmaxf(best, xr[c]);
      const float s_best = P.is_logit ? sigmoid_f32(best) : best;
      int cls = 0;
      for (int c = 0; c < P.C; ++c) {
        const float s = P.is_logit ? (xr[c] == best ? s_best : sigmoid_f32(xr[c])) : xr[c];
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
      float4 bx = P.boxes ? P.boxes[(size_t)b * P.n] : decode_box(P.deltas[(size_t)b * P.n], P.anchors[0], P.dp);
      P.out_boxes[o] = clip01(bx);
      P.out_scores[o] = -1.0f;
      P.out_classes[o] = -1;
    }
  }
}

// ===============================================================================================================
// K7  EfficientNMS_TRT-compatible entry (the node the reference appends for export mode onnx_tensorrt,
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
__global__ void topk_gather_per_class_kernel(const u64* __restrict__ emit_key /*[B*C][k]*/, const float4* __restrict__ boxes,
                                             int B, long n, int C, long k, float* __restrict__ scores_out,
                                             float4* __restrict__ boxes_out, int* __restrict__ idx_out) {
  const size_t tot = (size_t)B * k * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t bj = e / C;
    const long j = (long)(bj % k);
    const int b = (int)(bj / k);
    const u64 key = emit_key[((size_t)b * C + c) * k + j];
    const u32 row = key_tie(key);
    scores_out[e] = key_score(key);
    boxes_out[e] = boxes[(size_t)b * n + row];
    if (idx_out) idx_out[((size_t)b * C + c) * k + j] = (int)row;
  }
}

// global: emitted keys over the flat [n*C] axis; scores_out [B,k,C] = whole rows, boxes_out [B,k,4]
__global__ void topk_gather_global_kernel(const u64* __restrict__ emit_key /*[B][k]*/, const float* __restrict__ scores,
                                          const float4* __restrict__ boxes, int B, long n, int C, long k,
                                          float* __restrict__ scores_out, float4* __restrict__ boxes_out,
                                          int* __restrict__ idx_out) {
  const size_t tot = (size_t)B * k * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t bj = e / C;
    const int b = (int)(bj / k);
    const u32 flat = key_tie(emit_key[bj]);
    const u32 a = flat / (u32)C;  // indices // num_classes (:156)
    scores_out[e] = scores[((size_t)b * n + a) * C + c];
    if (c == 0) {
      boxes_out[bj] = boxes[(size_t)b * n + a];
      if (idx_out) idx_out[bj] = (int)flat;
    }
  }
}

// fused global filter: rows selected on raw logits -> materialise the reference's intermediates
// scores [B,k,C] = sigmoid(logit rows), boxes [B,k,4] = decoded anchors (TransformBoxesAndScores on k rows only)
__global__ void fused_global_rows_kernel(const u64* __restrict__ emit_key /*[B][k]*/, const float* __restrict__ logits,
                                         const float4* __restrict__ deltas, const float4* __restrict__ anchors,
                                         DecodeParams dp, int B, long N, int C, long k, int apply_sigmoid,
                                         float* __restrict__ scores_out, float4* __restrict__ boxes_out) {
  if (C < 16) {   // narrow rows: one thread per element
    const size_t tot = (size_t)B * k * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
      const int c = (int)(e % C);
      const size_t bj = e / C;
      const int b = (int)(bj / k);
      const u32 a = key_tie(emit_key[bj]) / (u32)C;
      const float raw = logits[((size_t)b * N + a) * C + c];
      scores_out[e] = apply_sigmoid ? sigmoid_f32(raw) : raw;
      if (c == 0) boxes_out[bj] = decode_box(deltas[(size_t)b * N + a], anchors[a], dp);
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
    const float* src = logits + ((size_t)b * N + a) * C;
    float* dst = scores_out + bj * C;
    // apply_sigmoid = 0 (Global* modes): the rows stay logits; only the row maxima are scored (global_pipeline)
    for (int c = lane; c < C; c += tpr) {
      const float raw = __ldg(src + c);
      dst[c] = apply_sigmoid ? sigmoid_f32(raw) : raw;
    }
    if (lane == 0) boxes_out[bj] = decode_box(deltas[(size_t)b * N + a], anchors[a], dp);
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
