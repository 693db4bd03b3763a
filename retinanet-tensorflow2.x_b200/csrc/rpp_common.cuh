// rpp_common.cuh — arithmetic and block primitives shared by the retinapost kernels (sm_100a).
//
// Arithmetic that decides results (decode, IoU, sigmoid) is written with explicit round-to-nearest intrinsics so
// that no FMA contraction can change a result relative to the reference's unfused TensorFlow ops
// (SURVEY.md A.1, A.6); the translation unit is additionally compiled with -fmad=false.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define RPP_FULL_MASK 0xffffffffu

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor has completed
// and its memory is visible (a no-op for a normal launch), pdl_launch() lets the successor's blocks be scheduled as
// soon as SM resources free up.  Every kernel of a fused pipeline calls both first thing, so stream order is kept and
// only launch latency, block scheduling and kernel prologues overlap the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch(); }

// ---------------------------------------------------------------------------------------------------------------
// keys: a candidate is ordered by (score descending, tie index ascending) — the total order of TF's TopKV2
//  and of NonMaxSuppressionV5's priority queue (A.2).  One u64, larger = better, 0 = invalid.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 ord_f32(float f) {
  f = __fadd_rn(f, 0.0f);  // -0.0 -> +0.0 (they compare equal in the reference)
  u32 b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(u32 o) {
  u32 b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
__device__ __forceinline__ u64 make_key(float score, u32 tie) {
  return ((u64)ord_f32(score) << 32) | (u64)(0xffffffffu - tie);
}
__device__ __forceinline__ float key_score(u64 k) { return unord_f32((u32)(k >> 32)); }
__device__ __forceinline__ u32 key_tie(u64 k) { return 0xffffffffu - (u32)k; }

// ---------------------------------------------------------------------------------------------------------------
// elementwise numerics
// ---------------------------------------------------------------------------------------------------------------
// tf.nn.sigmoid (postprocessing_ops.py:114): the exact logistic evaluated in binary64 and rounded once to
// binary32 (DESIGN.md "Numerics").  Only evaluated for candidates that survive the raw-logit pre-threshold.
// (not inlined: the binary64 exp is ~150 instructions and the problem kernels call it from a dozen sites; measured
// -7 us on the NMS stage of configs[1].  decode_box stays inlined: as a call it costs the soft per-class kernel 25 %.)
__device__ __noinline__ float sigmoid_f32(float x) { return (float)(1.0 / (1.0 + exp(-(double)x))); }
// tf.math.exp (postprocessing_ops.py:97)
__device__ __noinline__ float exp_f32(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float4 clip01(float4 b) {
  return make_float4(clip01(b.x), clip01(b.y), clip01(b.z), clip01(b.w));
}

struct DecodeParams {
  float shape[4];   // [H, W, H, W] applied to [x1, y1, x2, y2] (postprocessing_ops.py:65-69, :104)
  float var[4];     // encoder_params.box_variance
  int scale;        // encoder_params.scale_box_targets (:90-91)
};

// TransformBoxesAndScores._transform_box_predictions (postprocessing_ops.py:87-105); anchor = [cx, cy, w, h].
__device__ __forceinline__ float4 decode_box(float4 d, float4 a, const DecodeParams& p) {
  if (p.scale) {
    d.x = __fmul_rn(d.x, p.var[0]); d.y = __fmul_rn(d.y, p.var[1]);
    d.z = __fmul_rn(d.z, p.var[2]); d.w = __fmul_rn(d.w, p.var[3]);
  }
  const float cx = __fadd_rn(__fmul_rn(d.x, a.z), a.x);
  const float cy = __fadd_rn(__fmul_rn(d.y, a.w), a.y);
  const float hw = __fdiv_rn(__fmul_rn(exp_f32(d.z), a.z), 2.0f);
  const float hh = __fdiv_rn(__fmul_rn(exp_f32(d.w), a.w), 2.0f);
  return make_float4(__fdiv_rn(__fsub_rn(cx, hw), p.shape[0]), __fdiv_rn(__fsub_rn(cy, hh), p.shape[1]),
                     __fdiv_rn(__fadd_rn(cx, hw), p.shape[2]), __fdiv_rn(__fadd_rn(cy, hh), p.shape[3]));
}

// Boxes are canonicalised once for the IoU of TF's NMS kernels (iou_gt / iou_val in rpp_kernels.cuh):
// canon_box returns (min0, min1, max0, max1) and the area.
__device__ __forceinline__ float4 canon_box(float4 b, float& area) {
  float4 c = make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
  area = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
  return c;
}
// ---------------------------------------------------------------------------------------------------------------
// block primitives (NT threads, NT a multiple of 32, <= 1024)
// ---------------------------------------------------------------------------------------------------------------
template <int NT>
struct BlockScratch {
  u64 a[NT / 32];
  u64 b[NT / 32];
  u32 c[NT / 32];
  u64 ra, rb;
  u32 rc;
};

// count / max / min of a per-thread (cnt, mx, mn) triple; result broadcast to all threads.
template <int NT>
__device__ __forceinline__ void block_cnt_max_min(u32& cnt, u64& mx, u64& mn, BlockScratch<NT>* s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(RPP_FULL_MASK, cnt, o);
    u64 m2 = __shfl_xor_sync(RPP_FULL_MASK, mx, o);
    u64 n2 = __shfl_xor_sync(RPP_FULL_MASK, mn, o);
    mx = m2 > mx ? m2 : mx;
    mn = n2 < mn ? n2 : mn;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect scratch reuse
  if (l == 0) { s->a[w] = mx; s->b[w] = mn; s->c[w] = cnt; }
  __syncthreads();
  if (w == 0) {
    u64 m = l < NT / 32 ? s->a[l] : 0ull;
    u64 n = l < NT / 32 ? s->b[l] : ~0ull;
    u32 c = l < NT / 32 ? s->c[l] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_xor_sync(RPP_FULL_MASK, c, o);
      u64 m2 = __shfl_xor_sync(RPP_FULL_MASK, m, o);
      u64 n2 = __shfl_xor_sync(RPP_FULL_MASK, n, o);
      m = m2 > m ? m2 : m;
      n = n2 < n ? n2 : n;
    }
    if (l == 0) { s->ra = m; s->rb = n; s->rc = c; }
  }
  __syncthreads();
  mx = s->ra; mn = s->rb; cnt = s->rc;
}

// Descending bitonic sort of P2 (power of two) u64 keys in shared (or global) memory.
template <int NT>
__device__ __forceinline__ void bitonic_sort_desc(u64* k, int P2) {
  for (int size = 2; size <= P2; size <<= 1) {
    for (int j = size >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P2 >> 1); t += NT) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const u64 a = k[i], b = k[p];
        const bool desc = (i & size) == 0;
        if (desc ? (a < b) : (a > b)) { k[i] = b; k[p] = a; }
      }
      __syncthreads();
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Descending sort of exactly 8192 u64 keys in shared memory by a 1024-thread block: the same bitonic network as
// bitonic_sort_desc, but a thread keeps 8 keys in registers and most compare-exchange steps never touch shared memory.
// An element index has 13 bits.  Two register layouts:
//   A: bits 5-7 = register, bits 0-4 = lane, bits 8-12 = warp   -> steps j = 32..128 in registers, j = 1..16 by shuffle
//   C: bits 8-10 = register, bits 0-2 + 11-12 = lane, bits 3-7 = warp -> j = 256..1024 in registers, 2048 / 4096 by shuffle
// A merge phase of size >= 512 starts in C and finishes in A; the layouts are exchanged through the array itself
// (10 exchanges instead of 91 shared-memory passes with a barrier each).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cx_keep(u64& mine, u64 other, bool take_max) {
  const bool other_gt = other > mine;
  if (take_max == other_gt) mine = other;
}

template <bool LAYOUT_C>
__device__ __forceinline__ int sort8k_elem(int r, int lane, int warp) {
  return LAYOUT_C ? ((lane & 7) | (warp << 3) | (r << 8) | ((lane >> 3) << 11)) : ((warp << 8) | (r << 5) | lane);
}

template <bool LAYOUT_C, int SIZE, int J>
__device__ __forceinline__ void sort8k_step(u64 (&k)[8], int lane, int warp) {
  constexpr int REG_LO = LAYOUT_C ? 256 : 32;          // smallest stride held in registers
  if (J >= REG_LO && J < REG_LO * 8) {
    constexpr int m = (J / REG_LO) & 7;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r & m) continue;
      const int e = sort8k_elem<LAYOUT_C>(r, lane, warp);
      const bool desc = (e & SIZE) == 0;
      const u64 a = k[r], b = k[r | m];
      if (desc ? (a < b) : (a > b)) { k[r] = b; k[r | m] = a; }
    }
  } else {
    constexpr int lx = LAYOUT_C ? ((J >> 11) << 3) : J;   // lane bit of the partner
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const u64 other = __shfl_xor_sync(RPP_FULL_MASK, k[r], lx);
      const int e = sort8k_elem<LAYOUT_C>(r, lane, warp);
      const bool desc = (e & SIZE) == 0;
      const bool lower = (lane & lx) == 0;
      cx_keep(k[r], other, desc == lower);
    }
  }
}

template <bool LAYOUT_C>
__device__ __forceinline__ void sort8k_load(u64 (&k)[8], const u64* a, int lane, int warp) {
#pragma unroll
  for (int r = 0; r < 8; ++r) k[r] = a[sort8k_elem<LAYOUT_C>(r, lane, warp)];
}
template <bool LAYOUT_C>
__device__ __forceinline__ void sort8k_store(const u64 (&k)[8], u64* a, int lane, int warp) {
#pragma unroll
  for (int r = 0; r < 8; ++r) a[sort8k_elem<LAYOUT_C>(r, lane, warp)] = k[r];
}

template <int SIZE>
__device__ __forceinline__ void sort8k_phase_low(u64 (&k)[8], int lane, int warp) {   // steps j = min(SIZE/2, 128) .. 1
  if (SIZE > 128) sort8k_step<false, SIZE, 128>(k, lane, warp);
  if (SIZE > 64) sort8k_step<false, SIZE, 64>(k, lane, warp);
  if (SIZE > 32) sort8k_step<false, SIZE, 32>(k, lane, warp);
  if (SIZE > 16) sort8k_step<false, SIZE, 16>(k, lane, warp);
  if (SIZE > 8) sort8k_step<false, SIZE, 8>(k, lane, warp);
  if (SIZE > 4) sort8k_step<false, SIZE, 4>(k, lane, warp);
  if (SIZE > 2) sort8k_step<false, SIZE, 2>(k, lane, warp);
  sort8k_step<false, SIZE, 1>(k, lane, warp);
}

template <int SIZE>
__device__ __forceinline__ void sort8k_phase_high(u64 (&k)[8], u64* a, int lane, int warp) {   // SIZE >= 512
  __syncthreads();
  sort8k_store<false>(k, a, lane, warp);
  __syncthreads();
  sort8k_load<true>(k, a, lane, warp);
  if (SIZE > 4096) sort8k_step<true, SIZE, 4096>(k, lane, warp);
  if (SIZE > 2048) sort8k_step<true, SIZE, 2048>(k, lane, warp);
  if (SIZE > 1024) sort8k_step<true, SIZE, 1024>(k, lane, warp);
  if (SIZE > 512) sort8k_step<true, SIZE, 512>(k, lane, warp);
  sort8k_step<true, SIZE, 256>(k, lane, warp);
  __syncthreads();
  sort8k_store<true>(k, a, lane, warp);
  __syncthreads();
  sort8k_load<false>(k, a, lane, warp);
  sort8k_phase_low<SIZE>(k, lane, warp);
}

// All 1024 threads call; a[0 .. 8192) valid (pad with 0 = "no key": zeros sink to the end).
__device__ __noinline__ void block_sort8k_desc(u64* a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 k[8];
  sort8k_load<false>(k, a, lane, warp);
  sort8k_phase_low<2>(k, lane, warp);
  sort8k_phase_low<4>(k, lane, warp);
  sort8k_phase_low<8>(k, lane, warp);
  sort8k_phase_low<16>(k, lane, warp);
  sort8k_phase_low<32>(k, lane, warp);
  sort8k_phase_low<64>(k, lane, warp);
  sort8k_phase_low<128>(k, lane, warp);
  sort8k_phase_low<256>(k, lane, warp);
  sort8k_phase_high<512>(k, a, lane, warp);
  sort8k_phase_high<1024>(k, a, lane, warp);
  sort8k_phase_high<2048>(k, a, lane, warp);
  sort8k_phase_high<4096>(k, a, lane, warp);
  sort8k_phase_high<8192>(k, a, lane, warp);
  __syncthreads();
  sort8k_store<false>(k, a, lane, warp);
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
