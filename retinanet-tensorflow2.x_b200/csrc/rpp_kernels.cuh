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
// (no __restrict__: global_pipeline scores the row maxima in place, y == x)
__global__ void sigmoid_kernel(const float* x, float* y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = sigmoid_f32(x[i]);
}

__global__ void decode_kernel(const float4* __restrict__ deltas, const float4* __restrict__ anchors, long B, long N,
                              DecodeParams dp, float4* __restrict__ out) {
  const size_t tot = (size_t)B * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
    out[i] = decode_box(deltas[i], anchors[i % N], dp);
}

// The kernels proper, by stage (each file opens with the description of its stage):
#include "rpp_sample.cuh"    // K1  sample -> thresholds
#include "rpp_collect.cuh"   // K2  collect (the HBM-bound stream)
#include "rpp_nms.cuh"       // K3  problems: selection, NMS consumers, probe / bound, top-k emission
#include "rpp_outputs.cuh"   // K4-K8  merges, Global* outputs, gathers, EfficientNMS entry, COCO epilogue
#include "rpp_global.cuh"    // K9  Global* modes behind the global filter: row resolution, block-parallel soft NMS
