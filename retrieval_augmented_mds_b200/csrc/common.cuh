// common.cuh — shared definitions for the MIPS kernels (bank layout, top-k list helper).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

// Bank layout in HBM (one shard per GPU):
//   bank    [capacity, d_pad]  row-major, bf16 or fp32, d_pad = round_up(d, 64), columns
//                              [d, d_pad) are zero so kernels never need a K tail.
//   norm2   [capacity]         fp32 |x|^2 of the STORED (rounded / normalised) row.
// capacity is a multiple of 128 rows so a 128-row TMA box never leaves the allocation.
constexpr int kRowAlign = 128;
constexpr int kDimAlign = 64;

__host__ __device__ inline int round_up_i(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int64_t round_up_l(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Sorted (descending key, ascending id among equal keys) top-k list living in shared memory.
// Precondition: s > keys[k-1]. Candidates reach a list in ascending id order, so inserting
// behind every entry with key >= s keeps the (key desc, id asc) order. Returns the new k-th key.
__device__ __noinline__ float topk_list_insert(float* keys, int* ids, int k, float s, int id) {
  int p = k - 1;
  while (p > 0 && keys[p - 1] < s) {
    keys[p] = keys[p - 1];
    ids[p] = ids[p - 1];
    --p;
  }
  keys[p] = s;
  ids[p] = id;
  return keys[k - 1];
}

// ---------------------------------------------------------------------------------------------
// Per-thread top-k SETS (CTA-pair kernel, thread <-> query epilogue).
// Each query owns an UNSORTED set of KCAP >= k (ordered key, id) pairs in shared memory with the
// slot of its WORST entry tracked in a register. The 32 sets of a warp are INTERLEAVED: key i of
// lane t sits at warp_base[i * 32 + t] (ids follow the KCAP * 32 keys), so lanes that touch the same entry index together (the
// common case while many queries still admit candidates) hit 32 different banks; a per-lane
// contiguous layout made every such access a 32-way bank conflict (~2300 cycles per admission at
// k = 32, ncu: 43 % short-scoreboard stalls inside topk_replace). K2 merges sets by arg-max rounds and does not
// need them sorted. Keys are order-preserving uint32 images of the fp32 score; unused slots
// [k, KCAP) hold the never-worst sentinel 0xffffffff and are skipped on output.
__device__ __forceinline__ uint32_t f32_to_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}
__host__ __device__ inline int topk_kcap(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 32 ? 32 : 64; }

// Admission: store (sk, id) into the worst slot, then find the new worst entry: the lowest key and,
// among equal keys, the highest id (the last in the (key desc, id asc) order; empty slots carry id
// 0xffffffff). Keys and ids are separate arrays (`ids = keys + KCAP * 32`), so the common case only
// touches keys: the loads of a 32-key chunk are issued back to back, one integer min per tree node
// finds the lowest key, one compare per key marks where it sits; ids are read only to break an
// exact tie. ~130 instructions at k = 32 where a tournament on the 64-bit composite (key, ~id) took
// ~300 (~1000 cycles per admission, ncu) and the sorted insert / dynamic rescan loops before it
// measured ~50 cycles per entry of dependent load-compare-select.
constexpr int TOPK_STRIDE = 32;   // lanes of a warp interleave their sets entry by entry
template <int KCAP>
__device__ __forceinline__ uint2 topk_replace_fixed(uint32_t* keys, int worst, uint32_t sk, int id) {
  uint32_t* ids = keys + KCAP * TOPK_STRIDE;
  keys[worst * TOPK_STRIDE] = sk;
  ids[worst * TOPK_STRIDE] = static_cast<uint32_t>(id);
  constexpr int CH = KCAP < 32 ? KCAP : 32;
  constexpr int NCH = KCAP / CH;
  uint32_t m = 0xffffffffu;
  uint32_t eq[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t kk[CH], t[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) t[i] = kk[i] = keys[(c * CH + i) * TOPK_STRIDE];
#pragma unroll
    for (int w = 1; w < CH; w <<= 1)
#pragma unroll
      for (int i = 0; i + w < CH; i += 2 * w) t[i] = min(t[i], t[i + w]);
    const uint32_t mc = t[0];
    uint32_t e = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) e |= (kk[i] == mc ? 1u : 0u) << i;
    eq[c] = e;
    if (c == 0) {
      m = mc;
    } else if (mc < m) {      // a later chunk holds a lower key: earlier marks are void
      m = mc;
#pragma unroll
      for (int j = 0; j < c; ++j) eq[j] = 0;
    } else if (mc > m) {
      eq[c] = 0;
    }
  }
  int pos = -1;
  uint32_t best_id = 0;
  bool tie = false;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (eq[c]) {
      tie = tie || pos >= 0 || (eq[c] & (eq[c] - 1)) != 0;
      if (pos < 0) pos = c * CH + __ffs(eq[c]) - 1;
    }
  }
  if (tie) {   // rare: several entries share the lowest key, evict the one with the highest id
    pos = -1;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t e = eq[c];
      while (e) {
        const int i = c * CH + __ffs(e) - 1;
        e &= e - 1;
        const uint32_t v = ids[i * TOPK_STRIDE];
        if (pos < 0 || v > best_id) {
          best_id = v;
          pos = i;
        }
      }
    }
  }
  return make_uint2(m, static_cast<uint32_t>(pos));
}
// returns (worst key after the admission = new threshold, its slot)
__device__ __noinline__ uint2 topk_replace(uint32_t* keys, int kcap, int worst, uint32_t sk, int id) {
  if (kcap == 8) return topk_replace_fixed<8>(keys, worst, sk, id);
  if (kcap == 16) return topk_replace_fixed<16>(keys, worst, sk, id);
  if (kcap == 32) return topk_replace_fixed<32>(keys, worst, sk, id);
  return topk_replace_fixed<64>(keys, worst, sk, id);
}
