// k0_rows.cuh — K0: row ingest kernel (bank append and query preparation).
//
// One pass over fp32 rows [n, d]: optional L2 normalisation, cast to the storage type,
// zero padding to d_pad, |row|^2 of the stored values, and a running max of the ORIGINAL
// |row|^2. Fuses what the reference does in three host passes over the bank in
// Mips.build_index (sotasum/mips.py:298-331: _map_norm :347-349, _map_normalize :358-361 ->
// faiss.normalize_L2 :524, get_phi :55-56) and, for queries, Mips._prepare_query (:368-375).
//
// HBM-bound: algorithmic bytes per row = 4*d read + sizeof(T)*d_pad + 4 written.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid: ceil(n_pad / 8) blocks of 256 threads (8 warps, one row per warp).
// Rows [n, n_pad) of `out` are zero-filled (query tiles are padded to 128 rows).
template <typename T>
__global__ void __launch_bounds__(256) ingest_rows_kernel(const float* x, int64_t n,
                                                          int64_t n_pad, int d, int d_pad,
                                                          int normalize, T* out,  // may alias x (in place)
                                                          float* __restrict__ norm2_out,
                                                          unsigned int* __restrict__ max_norm2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  T* dst = out + row * d_pad;
  if (row >= n) {
    for (int i = lane; i < d_pad; i += 32) dst[i] = from_f32<T>(0.f);
    if (norm2_out) {
      if (lane == 0) norm2_out[row] = 0.f;
    }
    return;
  }
  const float* src = x + row * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = src[i];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  // faiss.normalize_L2: x *= 1/sqrt(|x|^2), rows with zero norm are left untouched.
  const float scale = (normalize && ss > 0.f) ? (1.0f / sqrtf(ss)) : 1.0f;
  float ss_stored = 0.f;
  for (int i = lane; i < d_pad; i += 32) {
    float v = 0.f;
    if (i < d) v = src[i] * scale;
    const T t = from_f32<T>(v);
    dst[i] = t;
    const float r = to_f32<T>(t);
    ss_stored = fmaf(r, r, ss_stored);
  }
  ss_stored = warp_sum(ss_stored);
  if (lane == 0) {
    if (norm2_out) norm2_out[row] = ss_stored;
    // non-negative floats order like their bit patterns; the running max only ever grows, so a
    // stale read can only cause a redundant atomic, never a missed one
    if (max_norm2_bits && __float_as_uint(ss) > *reinterpret_cast<volatile unsigned int*>(max_norm2_bits))
      atomicMax(max_norm2_bits, __float_as_uint(ss));
  }
}

// Vectorised variant for the common shapes (d % 4 == 0, d <= 128 * NV, 16-byte aligned rows):
// one warp per row, the row is read ONCE with 16-byte loads and held in registers between the
// norm pass and the store pass (the scalar kernel reads it twice with 4-byte loads: measured
// 2.3 TB/s = 35 % of the copy peak on 10M x 768 fp32 -> bf16). Stores are 16 bytes (fp32) or
// 8 bytes (4 bf16) per lane. Same arithmetic order per element as the scalar kernel except the
// lane-to-element assignment of the sums (fp32 partial sums regroup; |x|^2 differs in the last bits).
template <typename T, int NV>
__global__ void __launch_bounds__(256) ingest_rows_vec_kernel(const float* x, int64_t n, int64_t n_pad,
                                                              int d, int d_pad, int normalize, T* out,
                                                              float* __restrict__ norm2_out,
                                                              unsigned int* __restrict__ max_norm2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  T* dst = out + row * d_pad;
  const int nv = d >> 2, nv_pad = d_pad >> 2;   // float4 groups in the row / in the padded row
  float4 r[NV];
  float ss = 0.f;
  if (row < n) {
    const float4* src = reinterpret_cast<const float4*>(x + row * d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      r[i] = c < nv ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      ss = fmaf(r[i].x, r[i].x, ss);
      ss = fmaf(r[i].y, r[i].y, ss);
      ss = fmaf(r[i].z, r[i].z, ss);
      ss = fmaf(r[i].w, r[i].w, ss);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ss = warp_sum(ss);
  const float scale = (normalize && ss > 0.f) ? (1.0f / sqrtf(ss)) : 1.0f;
  float ss_stored = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv_pad) {
      const T t0 = from_f32<T>(r[i].x * scale), t1 = from_f32<T>(r[i].y * scale);
      const T t2 = from_f32<T>(r[i].z * scale), t3 = from_f32<T>(r[i].w * scale);
      const float f0 = to_f32<T>(t0), f1 = to_f32<T>(t1), f2 = to_f32<T>(t2), f3 = to_f32<T>(t3);
      ss_stored = fmaf(f0, f0, ss_stored);
      ss_stored = fmaf(f1, f1, ss_stored);
      ss_stored = fmaf(f2, f2, ss_stored);
      ss_stored = fmaf(f3, f3, ss_stored);
      if (sizeof(T) == 4) {
        reinterpret_cast<float4*>(dst)[c] = make_float4(f0, f1, f2, f3);
      } else {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1), hi = __floats2bfloat162_rn(f2, f3);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(dst)[c] = pk;
      }
    }
  }
  ss_stored = warp_sum(ss_stored);
  if (lane == 0) {
    if (norm2_out) norm2_out[row] = row < n ? ss_stored : 0.f;
    // One atomic per row on a single address serialises in L2 (measured: the kernel sat at 35 % of
    // the copy peak whatever the load width). The running max only ever grows, so compare with a
    // plain read first: a stale value can only cause a redundant atomic, never a missed one.
    if (max_norm2_bits && row < n &&
        __float_as_uint(ss) > *reinterpret_cast<volatile unsigned int*>(max_norm2_bits))
      atomicMax(max_norm2_bits, __float_as_uint(ss));
  }
}

// Copy stored rows back as fp32 [n, d] (Mips.save / np_search support).
template <typename T>
__global__ void __launch_bounds__(256) reconstruct_rows_kernel(const T* __restrict__ bank,
                                                               int64_t row0, int64_t n, int d,
                                                               int d_pad, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * d) return;
  const int64_t r = idx / d;
  const int c = static_cast<int>(idx - r * d);
  out[idx] = to_f32<T>(bank[(row0 + r) * d_pad + c]);
}

// ignore ids (global int64) -> local int32 row or -1 (not in this shard).
__global__ void ignore_to_local_kernel(const int64_t* __restrict__ ignore_ids, int nq,
                                       int64_t id_offset, int64_t ntotal, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int64_t l = ignore_ids[i] - id_offset;
  out[i] = (l >= 0 && l < ntotal) ? static_cast<int>(l) : -1;
}

// Multi-pass search (k > MIPS_MAX_K): pass p+1 only admits rows strictly AFTER the last result of pass p in the
// total order (key descending, id ascending). The bound id (global int64) becomes a shard-local row bound: rows
// with the bound's key are excluded iff row <= bound_row; ids of other shards clamp to -1 / INT_MAX.
__global__ void bound_to_local_kernel(const int64_t* __restrict__ after_id, int nq, int64_t id_offset,
                                      int* __restrict__ out_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int64_t l = after_id[i] - id_offset;
  out_row[i] = l < 0 ? -1 : (l > 0x7fffffffll ? 0x7fffffff : static_cast<int>(l));
}

// last column of a [nq, k] result -> the bound of the next pass (a query that ran out of rows, id -1, gets a bound
// nothing can follow)
__global__ void take_last_kernel(const float* __restrict__ key, const int64_t* __restrict__ ids, int nq, int k,
                                 float* __restrict__ after_key, int64_t* __restrict__ after_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int64_t id = ids[static_cast<size_t>(i) * k + (k - 1)];
  after_key[i] = id < 0 ? -CUDART_INF_F : key[static_cast<size_t>(i) * k + (k - 1)];
  after_id[i] = id < 0 ? INT64_MAX : id;
}
