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
// slot of its WORST entry tracked in a register; K2 merges sets by arg-max rounds and does not
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

// Admission: store (sk, id) into the worst slot, then find the new worst entry = the minimum of
// the composite (key, ~id), i.e. the lowest key and among equal keys the highest id (the last in
// the (key desc, id asc) order; empty slots carry id 0xffffffff). Straight-line code: the loads
// of a 16-entry chunk are issued back to back and reduced by a tournament, so an admission costs
// ~100 cycles per chunk; the sorted insert / dynamic rescan loops it replaces measured 800-1600
// cycles at k = 32 (ncu source page: ~50 cycles per entry of dependent load-compare-select).
template <int KCAP>
__device__ __forceinline__ uint2 topk_replace_fixed(uint2* set, int worst, uint32_t sk, int id) {
  set[worst] = make_uint2(sk, static_cast<uint32_t>(id));
  constexpr int CH = KCAP < 16 ? KCAP : 16;
  unsigned long long best = ~0ull;
  int best_pos = 0;
#pragma unroll
  for (int base = 0; base < KCAP; base += CH) {
    unsigned long long c[CH];
    int p[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const uint2 e = set[base + i];
      c[i] = (static_cast<unsigned long long>(e.x) << 32) | static_cast<uint32_t>(~e.y);
      p[i] = base + i;
    }
#pragma unroll
    for (int w = 1; w < CH; w <<= 1) {
#pragma unroll
      for (int i = 0; i + w < CH; i += 2 * w) {
        const bool lt = c[i + w] < c[i];
        c[i] = lt ? c[i + w] : c[i];
        p[i] = lt ? p[i + w] : p[i];
      }
    }
    if (c[0] < best) {
      best = c[0];
      best_pos = p[0];
    }
  }
  return make_uint2(static_cast<uint32_t>(best >> 32), static_cast<uint32_t>(best_pos));
}
// returns (worst key after the admission = new threshold, its slot)
__device__ __noinline__ uint2 topk_replace(uint2* set, int kcap, int worst, uint32_t sk, int id) {
  if (kcap == 8) return topk_replace_fixed<8>(set, worst, sk, id);
  if (kcap == 16) return topk_replace_fixed<16>(set, worst, sk, id);
  if (kcap == 32) return topk_replace_fixed<32>(set, worst, sk, id);
  return topk_replace_fixed<64>(set, worst, sk, id);
}
