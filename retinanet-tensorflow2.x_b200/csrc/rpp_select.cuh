// rpp_select.cuh — lazy, exact "next best chunk" selection over a keyed domain.
//
// Every consumer on this path (hard NMS, soft NMS, top-k emission, the per-image merge) reads candidates in the
// reference's total order (score desc, index asc) and usually stops long before the domain is exhausted (NMS
// stops at max_detections kept).  Instead of sorting a whole candidate list, a block repeatedly asks for the next
// chunk: the largest keys strictly below the running bound KB, at least ~`want` and at most CC of them, sorted
// descending in shared memory.  Keys are unique u64 (score bits | inverted tie index), so an MSB-first radix
// select on the key always terminates with an exact cut.  Works on any domain given by a functor key(i) -> u64
// (0 = not a candidate): a candidate list in shared or global memory, a column of the score tensor scanned
// directly (the exact slow path), or the C*M per-class results of the merge.
#pragma once
#include "rpp_common.cuh"

#ifndef RPP_RADIX_BITS
#define RPP_RADIX_BITS 10
#endif
#define RPP_RADIX_BINS (1 << RPP_RADIX_BITS)

template <int NT>
struct SelectScratch {
  BlockScratch<NT> bs;
  u32 hist[RPP_RADIX_BINS];
  u32 part[NT];       // per-thread bin sums for the suffix scan
  u32 d_star, cum_above, cnt_bin;
  u32 m;
};

// Returns m in [0, CC]: chunk[0..m) = the m largest keys of {key(i) : 0 < key(i) < KB}, sorted descending; KB is
// lowered to the smallest key returned.  m == 0 <=> the domain holds no key below KB.  m >= min(want/4, remaining).
// All threads of the block must call; chunk must hold next_pow2(CC) keys (CC keys when sort == false).
#ifndef RPP_SELECT_UNROLL
#define RPP_SELECT_UNROLL 8   // independent key loads in flight per thread (the sources are L2 / HBM resident lists and columns)
#endif
template <int NT, class KeyFn>
__device__ int select_chunk(KeyFn key, int n, u64& KB, int want, u64* chunk, int CC, SelectScratch<NT>* sc,
                            bool sort = true, u32* population = nullptr) {
  const int tid = threadIdx.x;
  // pass 1: population below the bound
  u32 cnt = 0;
  u64 mx = 0ull, mn = ~0ull;
  constexpr int U = RPP_SELECT_UNROLL;
  for (int i0 = tid; i0 < n; i0 += U * NT) {   // U independent key loads in flight per thread
    u64 k4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) k4[u] = i0 + u * NT < n ? key(i0 + u * NT) : 0ull;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const u64 k = k4[u];
      if (k != 0ull && k < KB) { ++cnt; mx = k > mx ? k : mx; mn = k < mn ? k : mn; }
    }
  }
  block_cnt_max_min<NT>(cnt, mx, mn, &sc->bs);
  if (population) *population = cnt;   // number of keys below the bound (per-thread copy of a uniform value)
  if (cnt == 0) return 0;

  u64 lo = mn;
  if ((int)cnt > CC) {
    if (want > CC) want = CC;
    if (want < 1) want = 1;
    const int min_ok = want / 4 > 0 ? want / 4 : 1;
    // MSB-first radix select below the common prefix of [mn, mx]
    const int p = 63 - __clzll((long long)(mn ^ mx));  // highest differing bit (mn != mx since keys are unique)
    int top_shift = p + 1;                              // bits >= top_shift are common
    u64 base = top_shift >= 64 ? 0ull : ((mx >> top_shift) << top_shift);
    u32 taken = 0;
    u32 total = cnt;
    for (;;) {
      const int bits = top_shift < RPP_RADIX_BITS ? top_shift : RPP_RADIX_BITS;
      const int shift = top_shift - bits;
      const int nb = 1 << bits;
      for (int i = tid; i < nb; i += NT) sc->hist[i] = 0;
      __syncthreads();
      for (int i0 = tid; i0 < n; i0 += U * NT) {
        u64 k4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) k4[u] = i0 + u * NT < n ? key(i0 + u * NT) : 0ull;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const u64 k = k4[u];
          if (k != 0ull && k < KB && k >= base && (top_shift >= 64 || ((k - base) >> top_shift) == 0ull))
            atomicAdd(&sc->hist[(u32)((k - base) >> shift)], 1u);
        }
      }
      __syncthreads();
      // find d* = the highest bin whose suffix count reaches w_eff
      u32 w_eff = (u32)want > taken ? (u32)want - taken : 1u;
      if (w_eff > total) w_eff = total;
      const int bpt = (nb + NT - 1) / NT;  // bins per thread (contiguous)
      const int b0 = tid * bpt;
      u32 local = 0;
      for (int j = 0; j < bpt; ++j)
        if (b0 + j < nb) local += sc->hist[b0 + j];
      // suffix over threads
      u32 above = 0;  // sum of bins owned by higher threads
      {
        // inclusive suffix sum inside the warp, then add the totals of the higher warps
        u32 v = local;
        const int lane = tid & 31;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_down_sync(RPP_FULL_MASK, v, o);
          if (lane + o < 32) v += t;
        }
        if (lane == 0) sc->part[tid >> 5] = v;   // warp total (part[] is reused: NT/32 <= NT)
        __syncthreads();
        for (int w = (tid >> 5) + 1; w < NT / 32; ++w) above += sc->part[w];
        above += v - local;
      }
      if (above < w_eff && above + local >= w_eff) {
        u32 run = above;
        for (int j = bpt - 1; j >= 0; --j) {
          const int bin = b0 + j;
          if (bin >= nb) continue;
          const u32 h = sc->hist[bin];
          if (run + h >= w_eff) { sc->d_star = (u32)bin; sc->cum_above = run; sc->cnt_bin = h; break; }
          run += h;
        }
      }
      __syncthreads();
      const u32 d = sc->d_star, cum_above = sc->cum_above, cnt_bin = sc->cnt_bin;
      if (taken + cum_above + cnt_bin <= (u32)CC) { lo = base + ((u64)d << shift); break; }
      if (taken + cum_above >= (u32)min_ok) { lo = base + ((u64)(d + 1) << shift); break; }
      taken += cum_above;
      base += (u64)d << shift;
      top_shift = shift;
      total = cnt_bin;
      __syncthreads();  // hist is rewritten next round
    }
  }
  // pass 3: compaction of {lo <= key < KB}.  One shared-memory atomic per warp and load round (ballot + popc),
  // not one per key: a chunk of a few thousand keys otherwise serialises on the single counter.
  if (tid == 0) sc->m = 0;
  __syncthreads();
  {
    const int lane = tid & 31;
    const u32 lt = (1u << lane) - 1u;
    for (int base = 0; base < n; base += U * NT) {   // uniform trip count: every lane takes part in the ballots
      u64 k4[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * NT + tid;
        k4[u] = i < n ? key(i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const u64 k = k4[u];
        const bool hit = k != 0ull && k < KB && k >= lo;
        const u32 mask = __ballot_sync(RPP_FULL_MASK, hit);
        if (mask == 0u) continue;
        u32 slot0 = 0u;
        if (lane == __ffs(mask) - 1) slot0 = atomicAdd(&sc->m, (u32)__popc(mask));
        slot0 = __shfl_sync(RPP_FULL_MASK, slot0, __ffs(mask) - 1);
        if (hit) {
          const u32 slot = slot0 + (u32)__popc(mask & lt);
          if (slot < (u32)CC) chunk[slot] = k;
        }
      }
    }
  }
  __syncthreads();
  int m = (int)sc->m;
  if (m > CC) m = CC;  // cannot happen (the cut guarantees <= CC); defensive
  int P2 = next_pow2(m < 2 ? 2 : m);
  // large chunks of a 1024-thread block (top-k emission): the register / shuffle sort of exactly 8192 slots
  const bool big = sort && NT == 1024 && CC >= 8192 && P2 > 2048;
  if (big) P2 = 8192;
  for (int i = m + tid; i < P2; i += NT) chunk[i] = 0ull;
  __syncthreads();
  if (big) block_sort8k_desc(chunk);
  else if (sort) bitonic_sort_desc<NT>(chunk, P2);
  KB = lo;
  return m;
}
