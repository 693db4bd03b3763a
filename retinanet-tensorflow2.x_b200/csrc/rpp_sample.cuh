// rpp_sample.cuh — K1: sampled per-(image, class) pre-thresholds.
// Part of the retinapost kernel set; included by rpp_kernels.cuh (one translation unit: rpp_api.cu).
#pragma once
#include "rpp_kernels.cuh"

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
  if (rl >= lanes) return;
  const int G = lanes * RPP_GPT;
  float m[RPP_GPT];
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i) m[i] = -INFINITY;
  const float* base = lv.x[0] + (size_t)b * N * C + c;   // fused tensor (LEVELS == false)
  const unsigned short* hbase = reinterpret_cast<const unsigned short*>(lv.x[0]) + (size_t)b * N * C + c;
  const int dtype = lv.dtype;
  for (int r = split; r < rounds; r += nsplit) {
    float v[RPP_GPT];
    if (LEVELS) {
      // A round covers G * stride consecutive rows of the fused axis and r is uniform over the block: the level of
      // the round comes from a select chain over the table in the kernel parameters and stays in uniform registers,
      // and the loads look like the fused tensor's.  The few rounds that straddle a level boundary (4 of ~30 at
      // 640 x 640) are SKIPPED: the sample only steers speed, and a second code path (or a shared-memory table with a
      // per-thread level cursor, as before: 35 us against the fused kernel's 21 us, spills at the 32-register budget)
      // costs more than the slightly noisier estimate.
      const long row_first = (long)r * G * stride, row_last = ((long)r * G + G - 1) * stride;
      long off_lo = lv.off[0], off_hi = lv.off[1];
      const float* xl = lv.x[0];
#pragma unroll
      for (int i = 1; i < RPP_MAX_LEVELS; ++i)
        if (i < lv.L && row_first >= lv.off[i]) { off_lo = lv.off[i]; off_hi = lv.off[i + 1]; xl = lv.x[i]; }
      if (row_last >= off_hi) continue;
      // lbase[row * C] (in elements of the input type) is element (b, row - off_l, c) of the level tensor
      const long shift = ((long)b * (off_hi - off_lo) - off_lo) * C + c;
#pragma unroll
      for (int i = 0; i < RPP_GPT; ++i) {
        const long s = (long)r * G + rl + i * lanes;
        if (HALF) v[i] = half_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(xl) + shift + (size_t)(s * stride) * C), dtype);
        else v[i] = __ldg(xl + shift + (size_t)(s * stride) * C);
      }
    } else {
#pragma unroll
      for (int i = 0; i < RPP_GPT; ++i) {
        const long s = (long)r * G + rl + i * lanes;  // sampled row index; group = rl + i * lanes
        if (HALF) v[i] = half_bits_to_f32(__ldg(hbase + (size_t)(s * stride) * C), dtype);
        else v[i] = __ldg(base + (size_t)(s * stride) * C);
      }
    }
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) m[i] = fmaxf(m[i], v[i]);
  }
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i)
    atomicMax(&gm[((size_t)b * G + rl + i * lanes) * C + c], ord_f32(m[i]));
}

// Single-column variant (C == 1: the flat anchors x classes axis of the global filter and of the EfficientNMS entry):
// every sample is one 16-byte load = four consecutive elements, so the same number of sampled elements touches a
// quarter of the sectors (a strided sample of single floats fetches 32 bytes for every 4 it uses).  Any n and any
// 4-byte alignment: image b is sampled over the whole 128-bit words that lie inside it (the <= 3 elements at each
// end are left out of the SAMPLE only — the collect pass sees every element).
__global__ void __launch_bounds__(1024, 2)
sample_max_flat4_kernel(const float* __restrict__ x /*16-byte aligned; element `lead` is x[0][0]*/, int lead, long n,
                        int stride4, int lanes, int rounds4, u32* __restrict__ gm /*[B][G]*/) {
  const int b = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
  const int rl = threadIdx.x;
  if (rl >= lanes) return;
  const int G = lanes * RPP_GPT;
  float m[RPP_GPT];
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i) m[i] = -INFINITY;
  const long fb = ((long)lead + (long)b * n + 3) >> 2;              // first whole word of image b
  const long nf = (((long)lead + (long)(b + 1) * n) >> 2) - fb;     // whole words inside image b
  const float4* base = reinterpret_cast<const float4*>(x) + fb;
  for (int r = split; r < rounds4; r += nsplit) {
    float4 v[RPP_GPT];
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) {
      const long idx = ((long)r * G + rl + i * lanes) * stride4;
      v[i] = idx < nf ? __ldg(base + idx) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int i = 0; i < RPP_GPT; ++i) m[i] = fmaxf(m[i], fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
#pragma unroll
  for (int i = 0; i < RPP_GPT; ++i) atomicMax(&gm[(size_t)b * G + rl + i * lanes], ord_f32(m[i]));
}

#define RPP_RANK_CPB 8   // classes per block
// rank_lo >= 0: when the rank_lo-th smallest group maximum is already below T_min, so few elements of the column pass
// the score threshold that ALL of them fit in the candidate list (host: choose_plan): the threshold is T_min itself and
// the list is complete — no exact column scan can ever be needed for that problem (trained-detector inputs: a class
// with a few objects has some hundreds of anchors above the threshold, all of them overlapping).
__global__ void sample_rank_kernel(const u32* __restrict__ gm, int C, int G, int rank, int rank_lo, float T_min,
                                   float* __restrict__ T) {
  pdl_enter();
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
    if (less <= rank && rank < less + eq) {
      // (the complete-list rule: count the group maxima below T_min directly)
      int below = 0;
      if (rank_lo >= 0)
        for (int g = 0; g < G; ++g) below += unord_f32(s_gm[g * RPP_RANK_CPB + cc]) < T_min;
      T[(size_t)b * C + c0 + cc] = (rank_lo >= 0 && below > rank_lo) ? T_min : fmaxf(unord_f32(v), T_min);
    }
  }
}

// Same result with a 128-key register bitonic sort per (image, class): one warp per class (G <= 128).
__global__ void __launch_bounds__(RPP_RANK_CPB * 32)
sample_rank_sort_kernel(const u32* __restrict__ gm, int C, int G, int rank, int rank_lo, float T_min,
                        float* __restrict__ T) {
  pdl_enter();
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
  u32 ans = 0u, lo = 0xffffffffu;
#pragma unroll
  for (int sidx = 0; sidx < 4; ++sidx) {
    if (sidx == rs) ans = v[sidx];
    if (rank_lo >= 0 && sidx == (rank_lo >> 5)) lo = v[sidx];
  }
  // complete-list rule (see sample_rank_kernel): element rank_lo of the ascending order below T_min
  lo = __shfl_sync(RPP_FULL_MASK, lo, rank_lo >= 0 ? (rank_lo & 31) : 0);
  const bool complete = rank_lo >= 0 && unord_f32(lo) < T_min;
  if (lane == rl) T[(size_t)b * C + c0 + cc] = complete ? T_min : fmaxf(unord_f32(ans), T_min);
}

__global__ void fill_kernel(float* p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
