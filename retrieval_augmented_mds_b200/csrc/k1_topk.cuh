// k1_topk.cuh — the fused per-query top-k epilogue shared by the tcgen05 search kernels
// (thread <-> query: a thread holds NG x 32 scores of its query's accumulator row in registers).
// The [nq, N] score matrix of the reference's brute force (sotasum/mips.py:552-560) never exists.
#pragma once
#include "common.cuh"

// v[c / 32][c % 32] for a run-time column c without spilling the register array to local memory.
template <int NG>
__device__ __forceinline__ uint32_t pick_col(const uint32_t (&v)[NG][32], int c) {
  uint32_t r = 0;
  switch (c) {
#define MIPS_PICK(I)                                                  \
  case I: r = v[0][I]; break;                                         \
  case 32 + I: r = v[1][I]; break;                                    \
  case 64 + I: if (NG > 2) r = v[NG > 2 ? 2 : 0][I]; break;           \
  case 96 + I: if (NG > 2) r = v[NG > 2 ? 3 : 0][I]; break;
    MIPS_PICK(0) MIPS_PICK(1) MIPS_PICK(2) MIPS_PICK(3) MIPS_PICK(4) MIPS_PICK(5) MIPS_PICK(6) MIPS_PICK(7)
    MIPS_PICK(8) MIPS_PICK(9) MIPS_PICK(10) MIPS_PICK(11) MIPS_PICK(12) MIPS_PICK(13) MIPS_PICK(14) MIPS_PICK(15)
    MIPS_PICK(16) MIPS_PICK(17) MIPS_PICK(18) MIPS_PICK(19) MIPS_PICK(20) MIPS_PICK(21) MIPS_PICK(22) MIPS_PICK(23)
    MIPS_PICK(24) MIPS_PICK(25) MIPS_PICK(26) MIPS_PICK(27) MIPS_PICK(28) MIPS_PICK(29) MIPS_PICK(30) MIPS_PICK(31)
#undef MIPS_PICK
  }
  return r;
}

// One accumulator row (NG x 32 scores of this thread's query), already in registers: fold it into the
// thread's top-k set (shared memory, unsorted, worst slot tracked). Fast path: one max tree against
// the admission threshold; the set is only touched when a score beats it.
template <bool kL2, int NG>
__device__ __forceinline__ void fold_tile(uint32_t (&v)[NG][32], const float* xnorm2, int id0,
                                          int64_t ntotal, int ign, bool live, uint32_t* set, int k, int kcap,
                                          bool first, float& thr, int& worst) {
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
  float m = -CUDART_INF_F;
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[g][j]));
  if (__builtin_expect(m > thr && live, 0)) {
    // Slow path, kept SMALL on purpose (an unrolled compare-and-call per column made the kernel
    // ~95 KB and ncu showed 42 % of its stall samples on instruction fetch): two instructions
    // per column build a candidate bit mask, then a rolled loop visits the few set bits and
    // pulls each score out of its register through one switch.
    int filled = 0;
    if (first && id0 + 64 <= ntotal && (ign < id0 || ign >= id0 + k)) {
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
      thr = ordered_to_f32(r.x);
      worst = static_cast<int>(r.y);
      filled = k;
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
    unsigned long long todo = (static_cast<unsigned long long>(mk[1]) << 32) | mk[0];
    unsigned long long later = (static_cast<unsigned long long>(mk[3]) << 32) | mk[2];
    int base = 0;
    while (true) {
      if (todo == 0ull) {
        if (base != 0 || later == 0ull) break;
        todo = later;
        base = 64;
      }
      const int c = base + __ffsll(static_cast<long long>(todo)) - 1;
      todo &= todo - 1;
      const float s = __uint_as_float(pick_col<NG>(v, c));
      const int id = id0 + c;
      if (s > thr && id < ntotal && id != ign) {
        const uint2 r = topk_replace(set, kcap, worst, f32_to_ordered(s), id);
        thr = ordered_to_f32(r.x);
        worst = static_cast<int>(r.y);
      }
    }
  }
}

