// k1_topk.cuh — the fused per-query top-k epilogue shared by the tcgen05 search kernels
// (thread <-> query: a thread holds NG x 32 scores of its query's accumulator row in registers).
// The [nq, N] score matrix of the reference's brute force (sotasum/mips.py:552-560) never exists.
#pragma once
#include "common.cuh"

// One of 64 registers (two 32-column groups) by a run-time index c in [0, 64), BRANCH FREE: a 6-level select
// tree (63 SEL). The first version was a `switch (c)` over the 128 columns; the compiler turned it into a compare
// tree of branches and the lanes of a warp — each with its own c — walked it one distinct path after the other:
// ncu on the 250k-row x 1024-query x k=32 shard put 35 % of all stall samples on that switch (~2000 cycles per
// admission). The select tree is the same ~70 instructions for every lane, fully convergent.
__device__ __forceinline__ uint32_t pick_col64(const uint32_t (&lo)[32], const uint32_t (&hi)[32], int c) {
  uint32_t a[32];
  const bool b5 = (c & 32) != 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = b5 ? hi[i] : lo[i];
  const bool b4 = (c & 16) != 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = b4 ? a[16 + i] : a[i];
  const bool b3 = (c & 8) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = b3 ? a[8 + i] : a[i];
  const bool b2 = (c & 4) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = b2 ? a[4 + i] : a[i];
  const bool b1 = (c & 2) != 0;
  a[0] = b1 ? a[2] : a[0];
  a[1] = b1 ? a[3] : a[1];
  return (c & 1) ? a[1] : a[0];
}

// Pooled admission threshold across the S bank splits of one query (the S CTA pairs that scan disjoint row
// ranges for the same queries, all resident at once). A split's own threshold (its k-th best so far) only
// tightens like k / rows_seen, so every split admits ~k ln(n_split / k) candidates — on a small shard with a
// large k that, not the MMA, was the kernel (250k rows, k = 32: 0.72 ms against 0.2 ms of MMA). But the answer
// only needs the GLOBAL top-k: with m = ceil(k / S), every split publishes the m-th best key it has admitted;
// once all S splits have published, T = min over splits of those values has >= S * m >= k entries >= T in the
// union of the splits' sets (a split never evicts its own m best while k >= m), so a candidate strictly below T
// cannot be in the global top-k and is skipped before it costs an admission. T only uses values that were true
// at SOME time (published values only grow), so stale reads are safe; a split that never publishes (fewer than m
// rows) just leaves pooling off. Exact: a candidate EQUAL to T is still admitted (ids break the tie later).
struct PoolState {
  float floor;     // candidates <= floor are skipped: the key just below T
  float top[4];    // this split's best admitted keys, descending
  int m;           // 0: pooling off
  __device__ __forceinline__ void init(int m_) {
    floor = top[0] = top[1] = top[2] = top[3] = -CUDART_INF_F;
    m = m_;
  }
  __device__ __forceinline__ void note(float s) {   // keep top[] sorted, descending
    if (s > top[3]) {
      const bool g2 = s > top[2], g1 = s > top[1], g0 = s > top[0];
      top[3] = g2 ? top[2] : s;
      top[2] = g1 ? top[1] : (g2 ? s : top[2]);
      top[1] = g0 ? top[0] : (g1 ? s : top[1]);
      top[0] = g0 ? s : top[0];
    }
  }
  __device__ __forceinline__ float mth() const { return m == 1 ? top[0] : m == 2 ? top[1] : m == 3 ? top[2] : top[3]; }
};

// One accumulator row (NG x 32 scores of this thread's query), already in registers: fold it into the
// thread's top-k set (shared memory, unsorted, worst slot tracked). Fast path: one max tree against
// the admission threshold; the set is only touched when a score beats it.
template <bool kL2, int NG, bool kPool>
__device__ __forceinline__ void fold_tile(uint32_t (&v)[NG][32], const float* xnorm2, int id0,
                                          int64_t ntotal, int ign, bool live, uint32_t* set, int k, int kcap,
                                          bool first, float& thr, int& worst, PoolState& pool,
                                          bool bounded = false, float bound_key = 0.f, int bound_row = -1) {
  if (kL2) {
    // ranking key for L2: <q,x> - |x|^2/2 (same address for every lane: broadcast loads)
    const float4* xn = reinterpret_cast<const float4*>(xnorm2 + id0);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(xn + g * 8 + j);
        v[g][4 * j + 0] = __float_as_uint(__uint_as_float(v[g][4 * j + 0]) - 0.5f * t.x);
        v[g][4 * j + 1] = __float_as_uint(__uint_as_float(v[g][4 * j + 1]) - 0.5f * t.y);
        v[g][4 * j + 2] = __float_as_uint(__uint_as_float(v[g][4 * j + 2]) - 0.5f * t.z);
        v[g][4 * j + 3] = __float_as_uint(__uint_as_float(v[g][4 * j + 3]) - 0.5f * t.w);
      }
    }
  }
  if (bounded) {
    // multi-pass search (k > MIPS_MAX_K): only rows strictly after the previous pass's last result in the order
    // (key descending, row ascending) are eligible; the others leave the tile as -inf (never admitted)
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(v[g][j]);
        const bool drop = s > bound_key || (s == bound_key && id0 + g * 32 + j <= bound_row);
        v[g][j] = drop ? __float_as_uint(-CUDART_INF_F) : v[g][j];
      }
  }
  float m = -CUDART_INF_F;
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[g][j]));
  if (__builtin_expect(m > thr && live, 0)) {
    // Slow path, kept SMALL on purpose (an unrolled compare-and-call per column made the kernel
    // ~95 KB and ncu showed 42 % of its stall samples on instruction fetch): two instructions
    // per column build a candidate bit mask, then a rolled loop visits the few set bits and
    // pulls each score out of its register through a branch-free select tree (pick_col64).
    int filled = 0;
    if (first && !bounded && id0 + 64 <= ntotal && (ign < id0 || ign >= id0 + k)) {
      // first tile of the split: its first k columns ARE the top-k so far. Store them directly (static
      // register indices, no admission calls: 128 admissions of ~300 cycles each otherwise open
      // every launch) and let one call find the worst entry.
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        if (i < k) {
          set[i * TOPK_STRIDE] = f32_to_ordered(__uint_as_float(v[i >> 5][i & 31]));
          set[(kcap + i) * TOPK_STRIDE] = static_cast<uint32_t>(id0 + i);
        }
      }
      const uint2 r = topk_replace(set, kcap, 0, f32_to_ordered(__uint_as_float(v[0][0])), id0);
      thr = kPool ? fmaxf(ordered_to_f32(r.x), pool.floor) : ordered_to_f32(r.x);
      worst = static_cast<int>(r.y);
      filled = k;
      if (kPool) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i < k) pool.note(__uint_as_float(v[i >> 5][i & 31]));
      }
    }
    uint32_t mk[4] = {0u, 0u, 0u, 0u};   // NG <= 4
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int j = 0; j < 32; ++j) mk[g] |= (__uint_as_float(v[g][j]) > thr ? 1u : 0u) << j;
    if (filled > 0) {   // columns [0, filled) are in the set already (filled <= 64)
      mk[0] &= filled >= 32 ? 0u : ~((1u << filled) - 1u);
      mk[1] &= filled <= 32 ? ~0u : (filled >= 64 ? 0u : ~((1u << (filled - 32)) - 1u));
    }
    // columns [0, 64) of every lane first, then [64, 128): ascending ids per query (the tie rule needs it), and
    // inside one loop the select tree spans two register groups only
#pragma unroll
    for (int h = 0; h < NG / 2; ++h) {
      unsigned long long todo = (static_cast<unsigned long long>(mk[2 * h + 1]) << 32) | mk[2 * h];
      while (todo != 0ull) {
        const int c = __ffsll(static_cast<long long>(todo)) - 1;
        todo &= todo - 1;
        const float s = __uint_as_float(pick_col64(v[2 * h], v[2 * h + 1], c));
        const int id = id0 + 64 * h + c;
        if (s > thr && id < ntotal && id != ign) {
          const uint2 r = topk_replace(set, kcap, worst, f32_to_ordered(s), id);
          thr = kPool ? fmaxf(ordered_to_f32(r.x), pool.floor) : ordered_to_f32(r.x);
          worst = static_cast<int>(r.y);
          if (kPool) pool.note(s);
        }
      }
    }
  }
}

