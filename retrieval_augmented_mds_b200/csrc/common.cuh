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
