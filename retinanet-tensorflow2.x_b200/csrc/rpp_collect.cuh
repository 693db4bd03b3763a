// rpp_collect.cuh — K2: the HBM-bound collect stream (fused, per-level, 16-bit and single-column variants).
// Part of the retinapost kernel set; included by rpp_kernels.cuh (one translation unit: rpp_api.cu).
#pragma once
#include "rpp_kernels.cuh"

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
  pdl_enter();
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
  pdl_enter();
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
  for (;;) {
    if (tid == 0) s_tile = (long)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
    if (tid == 0) s_span = 0u;
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const int b = (int)(tile / tiles_per_image);
    // Level geometry of the tile, by every thread from the block-uniform tile index with a select chain over the
    // table in the kernel parameters: the values stay in UNIFORM registers.  (Staged through shared memory by thread
    // 0 they came back in ordinary registers — four 64-bit values more per thread — and ptxas, at the 40-register
    // budget of 3 x 512 threads per SM, re-used the destinations of the four loads of a round, i.e. serialised them:
    // 282 us instead of 256 us for the same bytes.)
    const int t_img = (int)(tile % tiles_per_image);
    long off_lo = lv.off[0], off_hi = lv.off[1];
    int toff = 0;
    const float* xl = lv.x[0];
#pragma unroll
    for (int i = 1; i < RPP_MAX_LEVELS; ++i) {
      if (i < lv.L && t_img >= lv.tile_off[i]) { off_lo = lv.off[i]; off_hi = lv.off[i + 1]; toff = lv.tile_off[i]; xl = lv.x[i]; }
    }
    // rows are LOCAL to the level inside the loop; `goff` turns them into fused row indices when staged
    const long n_l = off_hi - off_lo, goff = off_lo, r0 = (long)(t_img - toff) * rows_per_tile;
    const float* __restrict__ xb = xl + (size_t)b * n_l * C;
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
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = C8 * 8;
  uint2* s_stage = reinterpret_cast<uint2*>(smem_raw);
  u32* s_cnt = reinterpret_cast<u32*>(s_stage + (size_t)C * RPP_STAGE_CAP);
  u32* s_base = s_cnt + C;
  __shared__ long s_tile;
  __shared__ u32 s_span;
  const int tid = threadIdx.x;
  const int co = tid % C8, rl = tid / C8;
  const bool active = rl < lanes;
  const int dtype = lv.dtype;
  const long n_tiles = (long)B * tiles_per_image;
  for (;;) {
    if (tid == 0) s_tile = (long)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
    if (tid == 0) s_span = 0u;
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const int b = (int)(tile / tiles_per_image);
    // level geometry from the block-uniform tile index (uniform registers: see collect_cols4_levels_kernel)
    const int t_img = (int)(tile % tiles_per_image);
    long off_lo = lv.off[0], off_hi = lv.off[1];
    int toff = 0;
    const float* xl = lv.x[0];
#pragma unroll
    for (int i = 1; i < RPP_MAX_LEVELS; ++i) {
      if (i < lv.L && t_img >= lv.tile_off[i]) { off_lo = lv.off[i]; off_hi = lv.off[i + 1]; toff = lv.tile_off[i]; xl = lv.x[i]; }
    }
    const long n_l = off_hi - off_lo, goff = off_lo, r0 = (long)(t_img - toff) * rows_per_tile;
    const long r1 = r0 + rows_per_tile < n_l ? r0 + rows_per_tile : n_l;
    const size_t pbase = (size_t)b * C;
    const unsigned short* __restrict__ xb = reinterpret_cast<const unsigned short*>(xl) + (size_t)b * n_l * C;
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
// Global* modes): x [B, n], one threshold per image, ANY n.  The batch is treated as one flat array of B * n floats
// (only its base must be 16-byte aligned): tiles are ranges of 128-bit words of that array, so image bases that are
// not multiples of 16 bytes (n % 4 != 0, e.g. 320x320 / 5 classes: n = 96 030; or a tensor whose own base is only
// 4-byte aligned: `lead`) still stream with LDG.128.  A tile that
// straddles an image boundary is processed as one segment per image; the <= 3 elements at each ragged segment edge
// are read with scalar loads, and no 128-bit load ever crosses a segment (or the end of the array).
// Hits are counted per thread (a 4 * UNROLL bit mask), placed in the shared queue with ONE shared atomic per warp and
// load round (warp prefix sum), and flushed with one global atomic per (tile, image).
#define RPP_FLAT_QCAP 2048
// element k of a 128-bit word as fp32: 4 floats, or 8 f16 / bf16 values (converted exactly)
template <int DT> struct FlatWord;
template <> struct FlatWord<RPP_DT_F32> {
  static constexpr int EPW = 4;
  static __device__ __forceinline__ float get(const uint4& w, int k) {
    return __uint_as_float(k == 0 ? w.x : k == 1 ? w.y : k == 2 ? w.z : w.w);
  }
  static __device__ __forceinline__ float load(const void* x, long e) { return __ldg(reinterpret_cast<const float*>(x) + e); }
};
template <int DT> struct FlatWordHalf {
  static constexpr int EPW = 8;
  static __device__ __forceinline__ float get(const uint4& w, int k) {
    const u32 v = (k >> 1) == 0 ? w.x : (k >> 1) == 1 ? w.y : (k >> 1) == 2 ? w.z : w.w;
    return half_bits_to_f32((unsigned short)((k & 1) ? (v >> 16) : (v & 0xffffu)), DT);
  }
  static __device__ __forceinline__ float load(const void* x, long e) {
    return half_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(x) + e), DT);
  }
};
template <> struct FlatWord<RPP_DT_F16> : FlatWordHalf<RPP_DT_F16> {};
template <> struct FlatWord<RPP_DT_BF16> : FlatWordHalf<RPP_DT_BF16> {};

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// x: 16-byte aligned address; element `lead` (0 .. EPW-1) of it is x[0][0].  idx_off: added to the stored indices (a
// per-level piece of a longer axis: rpp_detect_levels / rpp_detect_typed with the global filter).
template <int UNROLL, int DT>
__global__ void __launch_bounds__(RPP_COLLECT_NT, 3)
collect_flat_kernel(const void* __restrict__ x, int lead, u32 idx_off, const float* __restrict__ T /*[B]*/,
                    u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, int B, long n, long tile_elems,
                    long n_tiles, u32* __restrict__ tile_counter) {
  pdl_enter();
  typedef FlatWord<DT> W;
  constexpr int EPW = W::EPW;
  __shared__ uint2 s_q[RPP_FLAT_QCAP];
  __shared__ u32 s_qn, s_base;
  __shared__ long s_tile;
  const int tid = threadIdx.x, lane = tid & 31;
  const long total = (long)lead + (long)B * n;   // the flat array, counted from the aligned address
  const uint4* __restrict__ x4 = reinterpret_cast<const uint4*>(x);
  for (;;) {
    if (tid == 0) { s_tile = (long)atomicAdd(tile_counter, 1u); s_qn = 0u; }
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const long e0 = tile * tile_elems;
    const long e1 = e0 + tile_elems < total ? e0 + tile_elems : total;
    for (long b = (e0 > lead ? e0 - lead : 0) / n; b < B && lead + b * n < e1; ++b) {   // images of this tile
      const long ibase = lead + b * n;
      const long seg_lo = e0 > ibase ? e0 : ibase;
      const long seg_hi = e1 < ibase + n ? e1 : ibase + n;
      const float t = __ldg(T + b);
      long v_lo = (seg_lo + EPW - 1) / EPW * EPW, v_hi = seg_hi / EPW * EPW;   // whole 128-bit words inside the segment
      if (v_lo > v_hi) v_lo = v_hi = seg_hi;
      {   // ragged edges: [seg_lo, v_lo) and [v_hi, seg_hi), at most EPW-1 elements each
        const long head = (v_lo < seg_hi ? v_lo : seg_hi) - seg_lo;
        const long tail = seg_hi - (v_hi > v_lo ? v_hi : v_lo);
        if (tid < head + tail) {
          const long e = tid < head ? seg_lo + tid : (v_hi > v_lo ? v_hi : v_lo) + (tid - head);
          const float v = W::load(x, e);
          if (v >= t) {
            const u32 slot = atomicAdd(&s_qn, 1u);
            if (slot < RPP_FLAT_QCAP) s_q[slot] = make_uint2(__float_as_uint(v), (u32)(e - ibase) + idx_off);
            else append_cand(cand_count, cand, CAP, (size_t)b, v, (u32)(e - ibase) + idx_off);
          }
        }
      }
      const long f_lo = v_lo / EPW, f_hi = v_hi / EPW;
      for (long f0 = f_lo; f0 < f_hi; f0 += (long)RPP_COLLECT_NT * UNROLL) {   // uniform trip count per block
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long ff = f0 + tid + (long)u * RPP_COLLECT_NT;
          v[u] = ff < f_hi ? ld_stream_u4(x4 + ff) : make_uint4(0u, 0u, 0u, 0u);
        }
        u32 mask = 0u;   // bit EPW * u + i: element i of load u passes (NaN never passes >=)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const bool in = f0 + tid + (long)u * RPP_COLLECT_NT < f_hi;
#pragma unroll
          for (int i = 0; i < EPW; ++i)
            if (in && W::get(v[u], i) >= t) mask |= 1u << (EPW * u + i);
        }
        if (__any_sync(RPP_FULL_MASK, mask != 0u)) {
          // warp prefix sum of the per-thread hit counts -> one shared atomic per warp
          const u32 cnt = (u32)__popc(mask);
          u32 incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const u32 up = __shfl_up_sync(RPP_FULL_MASK, incl, o);
            if (lane >= o) incl += up;
          }
          u32 wbase = 0u;
          if (lane == 31) wbase = atomicAdd(&s_qn, incl);
          wbase = __shfl_sync(RPP_FULL_MASK, wbase, 31);
          u32 slot = wbase + incl - cnt;
          while (mask) {
            const int bit = __ffs(mask) - 1;
            mask &= mask - 1u;
            const long e = (f0 + tid + (long)(bit / EPW) * RPP_COLLECT_NT) * EPW + (bit % EPW);
            const u32 idx = (u32)(e - ibase) + idx_off;
            const float val = W::load(x, e);   // L1 hit: the line was just loaded by this thread
            if (slot < RPP_FLAT_QCAP) s_q[slot] = make_uint2(__float_as_uint(val), idx);
            else append_cand(cand_count, cand, CAP, (size_t)b, val, idx);
            ++slot;
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
      if (tid == 0) s_qn = 0u;
      __syncthreads();
    }
  }
}

// Any-C variant of the column collect (num_classes % 4 != 0, e.g. 91-class or 5-class heads; fp32, fused tensor):
// the batch is one flat array of 128-bit words again, and the word stride between a thread's consecutive loads is a
// multiple S of C / gcd(C, 4), so the classes of a thread's four components never change ("phase invariance"): its
// four thresholds stay in registers for a whole image segment, exactly as in collect_cols4_kernel.  Tiles start at
// multiples of S words; a tile that straddles an image boundary is processed one image segment at a time, the <= 3
// elements at each ragged segment edge with scalar loads.  Staging and flush as in collect_cols4_kernel.
template <int UNROLL>
__global__ void __launch_bounds__(RPP_COLLECT_NT, 3)
collect_colsv_kernel(const float* __restrict__ x /*[B*N*C], 16-byte aligned*/, const float* __restrict__ T /*[B*C]*/,
                     u32* __restrict__ cand_count, uint2* __restrict__ cand, int CAP, int B, long N, int C,
                     int S /*threads that load: 4 * S % C == 0*/, long tile_f4 /*multiple of S * UNROLL*/, long n_tiles,
                     u32* __restrict__ tile_counter) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint2* s_stage = reinterpret_cast<uint2*>(smem_raw);   // [C][RPP_STAGE_CAP]
  u32* s_cnt = reinterpret_cast<u32*>(s_stage + (size_t)C * RPP_STAGE_CAP);
  u32* s_base = s_cnt + C;
  __shared__ long s_tile;
  __shared__ u32 s_span;
  const int tid = threadIdx.x;
  const long NC = N * (long)C;
  const long total = (long)B * NC;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  int cls[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) cls[i] = (int)((4L * tid + i) % C);
  auto stage_hit = [&](size_t pbase, int c, float val, u32 row) {
    const u32 slot = atomicAdd(&s_cnt[c], 1u);
    if (slot < RPP_STAGE_CAP) s_stage[c * RPP_STAGE_CAP + slot] = make_uint2(__float_as_uint(val), row);
    else append_cand(cand_count, cand, CAP, pbase + c, val, row);
  };
  for (;;) {
    if (tid == 0) s_tile = (long)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
    if (tid == 0) s_span = 0u;
    __syncthreads();
    const long tile = s_tile;
    if (tile >= n_tiles) break;
    const long f0 = tile * tile_f4;
    const long e0 = 4 * f0;
    long e1 = e0 + 4 * tile_f4;
    if (e1 > total || tile == n_tiles - 1) e1 = total;          // (the last tile also takes the <= 3 tail elements)
    for (long b = e0 / NC; b < B && b * NC < e1; ++b) {
      const long ibase = b * NC;
      const long seg_lo = e0 > ibase ? e0 : ibase;
      const long seg_hi = e1 < ibase + NC ? e1 : ibase + NC;
      const size_t pbase = (size_t)b * C;
      long v_lo = (seg_lo + 3) & ~3L, v_hi = seg_hi & ~3L;
      if (v_lo > v_hi) v_lo = v_hi = seg_hi;
      {   // ragged edges
        const long head = (v_lo < seg_hi ? v_lo : seg_hi) - seg_lo;
        const long tail = seg_hi - (v_hi > v_lo ? v_hi : v_lo);
        if (tid < head + tail) {
          const long e = tid < head ? seg_lo + tid : (v_hi > v_lo ? v_hi : v_lo) + (tid - head);
          const float v = __ldg(x + e);
          const long eo = e - ibase;
          const int c = (int)(eo % C);
          if (v >= __ldg(T + pbase + c)) stage_hit(pbase, c, v, (u32)(eo / C));
        }
      }
      if (tid < S) {
        const float t0 = __ldg(T + pbase + cls[0]), t1 = __ldg(T + pbase + cls[1]);
        const float t2 = __ldg(T + pbase + cls[2]), t3 = __ldg(T + pbase + cls[3]);
        const long f_lo = v_lo >> 2, f_hi = v_hi >> 2;
        // this thread's words: f = f0 + tid + it * S (the phase is counted from the tile start)
        long f = f0 + tid;
        if (f < f_lo) f += (f_lo - f + S - 1) / S * S;
        for (; f < f_hi; f += (long)S * UNROLL) {
          float4 v[UNROLL];
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const long ff = f + (long)u * S;
            v[u] = ff < f_hi ? ld_stream_f4(x4 + ff) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          }
          u32 mask = 0u;
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
            mask |= ((v[u].x >= t0 ? 1u : 0u) | (v[u].y >= t1 ? 2u : 0u) | (v[u].z >= t2 ? 4u : 0u) |
                     (v[u].w >= t3 ? 8u : 0u)) << (4 * u);
          while (mask) {
            const int bit = __ffs(mask) - 1;
            mask &= mask - 1u;
            const long e = (f + (long)(bit >> 2) * S) * 4 + (bit & 3);
            const float val = __ldg(x + e);   // L1 hit: the line was just loaded by this thread
            stage_hit(pbase, cls[bit & 3], val, (u32)((e - ibase) / C));
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
      const int span = (int)s_span;
      for (int e = tid; e < C * span; e += RPP_COLLECT_NT) {
        const int c = e / span, r = e - c * span;
        const u32 n = s_cnt[c] < RPP_STAGE_CAP ? s_cnt[c] : RPP_STAGE_CAP;
        if ((u32)r < n) {
          const u32 slot = s_base[c] + (u32)r;
          if (slot < (u32)CAP) cand[(pbase + c) * (size_t)CAP + slot] = s_stage[c * RPP_STAGE_CAP + r];
        }
      }
      __syncthreads();
      for (int i = tid; i < C; i += RPP_COLLECT_NT) s_cnt[i] = 0u;
      if (tid == 0) s_span = 0u;
      __syncthreads();
    }
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
