// k3_rerank.cuh — exact fp32 search at tensor-core speed: bf16 shadow + exact re-rank + certificate.
//
// The fp32 index (BASELINE config 2: results must equal the reference's exact search, faiss
// IndexFlat / inner_product, sotasum/mips.py:383-386, :552-560) keeps, next to its fp32 rows, a
// bf16-rounded SHADOW of them. A search is then
//   1. K1 (tcgen05 pair kernel) over the shadow with the bf16-rounded queries, keeping kc > k
//      candidates per query (approximate keys a = fl(<qh, xh>) [- |x|^2/2]),
//   2. rerank_exact_kernel: the candidates' keys recomputed from the fp32 rows with fp32 FMAs
//      (what the SIMT kernel computes), top-k of those by (key desc, id asc),
//   3. a per-query CERTIFICATE that no row outside the candidate set can enter the top-k:
//        every outside row has a <= T = max(kc-th approximate key of the merge, the m-th best key of
//        every split that filled its m-entry list), and for every row
//        |a - <q,x>| <= E(q) = |ql| * Xh + |q| * XL + 2 g |q| Xh
//        (q = qh + ql, x = xh + xl; Xh >= max |xh|, XL = max |xl|, g = d_pad * 2^-22 covers the fp32
//        accumulation of the tensor pipe and of the re-rank), so S_k > T + E(q) proves exactness.
//   4. queries that fail the certificate (ties and near-ties at the boundary) are recomputed by
//      the exact SIMT kernel; the certificate is rigorous, so the result is always the exact one.
#pragma once
#include "common.cuh"
#include "k0_rows.cuh"
#include "k2_merge.cuh"

// bf16 shadow of fp32 rows [n, d_pad] (+ per-row |x - bf16(x)|^2 and running maxima of |x|^2 and of
// the residual). One warp per row, 16-byte loads, 8-byte stores.
__global__ void __launch_bounds__(256) shadow_rows_kernel(const float* __restrict__ src, int64_t n, int d_pad,
                                                          __nv_bfloat16* __restrict__ dst,
                                                          float* __restrict__ res2_out,          // [n] or null
                                                          unsigned int* __restrict__ max_norm2_bits,  // or null
                                                          unsigned int* __restrict__ max_res2_bits) { // or null
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float4* s4 = reinterpret_cast<const float4*>(src + row * d_pad);
  uint2* d2 = reinterpret_cast<uint2*>(dst + row * d_pad);
  float nn = 0.f, rr = 0.f;
  for (int c = lane; c < d_pad / 4; c += 32) {
    const float4 v = s4[c];
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    d2[c] = pk;
    const float r0 = v.x - __low2float(lo), r1 = v.y - __high2float(lo);
    const float r2 = v.z - __low2float(hi), r3 = v.w - __high2float(hi);
    nn = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, nn))));
    rr = fmaf(r0, r0, fmaf(r1, r1, fmaf(r2, r2, fmaf(r3, r3, rr))));
  }
  nn = warp_sum(nn);
  rr = warp_sum(rr);
  if (lane == 0) {
    if (res2_out) res2_out[row] = rr;
    if (max_norm2_bits && __float_as_uint(nn) > *reinterpret_cast<volatile unsigned int*>(max_norm2_bits))
      atomicMax(max_norm2_bits, __float_as_uint(nn));
    if (max_res2_bits && __float_as_uint(rr) > *reinterpret_cast<volatile unsigned int*>(max_res2_bits))
      atomicMax(max_res2_bits, __float_as_uint(rr));
  }
}

// One block (8 warps) per query: exact keys of its kc <= 128 candidates, top-k, certificate.
template <bool kL2>
__global__ void __launch_bounds__(256) rerank_exact_kernel(
    const float* __restrict__ q_prep,      // [nq_pad, d_pad] prepared fp32 queries
    const float* __restrict__ bank,        // [capacity, d_pad] fp32 rows
    const float* __restrict__ xnorm2,      // [capacity]
    int d_pad, const float* __restrict__ cand_key, const int64_t* __restrict__ cand_rows,  // [nq, kc]
    int nq, int kc, int k,
    const float* __restrict__ part_key, const int* __restrict__ part_ids, int n_parts, int m,  // [n_parts, nq, m]
    const float* __restrict__ q_norm2, const float* __restrict__ q_res2,
    const unsigned int* __restrict__ max_norm2_bits, const unsigned int* __restrict__ max_res2_bits,
    int64_t id_offset, float* __restrict__ out_key, int64_t* __restrict__ out_ids,
    float* __restrict__ out_xn2, PackedCand* __restrict__ out_packed,
    int* __restrict__ need_fallback,       // [nq]
    int* __restrict__ tile_flag,           // [ceil(nq / 64)] SIMT query tiles to recompute
    int* __restrict__ n_fallback) {        // device counter (statistics)
  constexpr int KC_MAX = 2 * MIPS_MAX_K;   // candidates per query
  constexpr int PER_LANE = KC_MAX / 32;
  __shared__ float s_exact[KC_MAX];
  __shared__ float s_approx[KC_MAX];
  __shared__ int s_row[KC_MAX];
  __shared__ float s_t0[8];
  const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // T0: a row that its split dropped has an approximate key <= that split's m-th best (the lowest
  // key of its FULL list; a list with free slots dropped nothing)
  {
    float t0 = -CUDART_INF_F;
    for (int p = threadIdx.x; p < n_parts; p += blockDim.x) {
      const size_t a = (static_cast<size_t>(p) * nq + q) * m;
      float lo = CUDART_INF_F;
      bool full = true;
      for (int i = 0; i < m; ++i) {
        full = full && part_ids[a + i] >= 0;
        lo = fminf(lo, part_key[a + i]);
      }
      if (full) t0 = fmaxf(t0, lo);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, o));
    if (lane == 0) s_t0[warp] = t0;
  }
  const float4* q4 = reinterpret_cast<const float4*>(q_prep + static_cast<size_t>(q) * d_pad);
  for (int c = warp; c < kc; c += 8) {
    const int64_t row = cand_rows[static_cast<size_t>(q) * kc + c];
    float key = -CUDART_INF_F;
    if (row >= 0) {
      const float4* x4 = reinterpret_cast<const float4*>(bank + static_cast<size_t>(row) * d_pad);
      float acc = 0.f;
      for (int i = lane; i < d_pad / 4; i += 32) {
        const float4 a = q4[i], b = __ldg(x4 + i);
        acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
      }
      acc = warp_sum(acc);
      key = kL2 ? acc - 0.5f * xnorm2[row] : acc;
    }
    if (lane == 0) {
      s_exact[c] = key;
      s_approx[c] = cand_key[static_cast<size_t>(q) * kc + c];
      s_row[c] = static_cast<int>(row);
    }
  }
  __syncthreads();
  if (warp != 0) return;

  // lane l owns candidates l, l + 32, l + 64, l + 96
  float ek[PER_LANE], ak[PER_LANE];
  int er[PER_LANE];
#pragma unroll
  for (int t = 0; t < PER_LANE; ++t) {
    const int c = lane + 32 * t;
    const bool in = c < kc && s_row[c] >= 0;
    ek[t] = in ? s_exact[c] : -CUDART_INF_F;
    ak[t] = in ? s_approx[c] : CUDART_INF_F;
    er[t] = in ? s_row[c] : -1;
  }
  int n_valid = 0;
#pragma unroll
  for (int t = 0; t < PER_LANE; ++t) n_valid += __popc(__ballot_sync(0xffffffffu, er[t] >= 0));
  // T1: a row that reached the merge but not the kc candidates has an approximate key <= the kc-th
  float t_min = CUDART_INF_F;
#pragma unroll
  for (int t = 0; t < PER_LANE; ++t) t_min = fminf(t_min, ak[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t_min = fminf(t_min, __shfl_xor_sync(0xffffffffu, t_min, o));
  float t_bound = n_valid == kc ? t_min : -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < 8; ++i) t_bound = fmaxf(t_bound, s_t0[i]);

  // k rounds of arg-max in the (key desc, row asc) order, each strictly after the previous pick
  float prev_key = CUDART_INF_F, kth = -CUDART_INF_F;
  int prev_row = -1, n_found = 0;
  for (int j = 0; j < k; ++j) {
    float bk = -CUDART_INF_F;
    int br = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < PER_LANE; ++t) {
      if (er[t] < 0) continue;
      const bool after = ek[t] < prev_key || (ek[t] == prev_key && er[t] > prev_row);
      if (after && (ek[t] > bk || (ek[t] == bk && er[t] < br))) {
        bk = ek[t];
        br = er[t];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int orow = __shfl_xor_sync(0xffffffffu, br, o);
      if (orow != 0x7fffffff && (br == 0x7fffffff || ok > bk || (ok == bk && orow < br))) {
        bk = ok;
        br = orow;
      }
    }
    const size_t o = static_cast<size_t>(q) * k + j;
    if (br == 0x7fffffff) {
      if (lane == 0) {
        if (out_packed) out_packed[o] = PackedCand{-CUDART_INF_F, 0.f, -1};
        else {
          out_key[o] = -CUDART_INF_F;
          out_ids[o] = -1;
          if (out_xn2) out_xn2[o] = 0.f;
        }
      }
      continue;
    }
    prev_key = bk;
    prev_row = br;
    kth = bk;
    n_found = j + 1;
    if (lane == 0) {
      const float xn = xnorm2[br];
      if (out_packed) out_packed[o] = PackedCand{bk, xn, id_offset + br};
      else {
        out_key[o] = bk;
        out_ids[o] = id_offset + br;
        if (out_xn2) out_xn2[o] = xn;
      }
    }
  }
  if (lane == 0) {
    const float qn = sqrtf(q_norm2[q]), ql = sqrtf(q_res2[q]);
    const float xl = sqrtf(__uint_as_float(*max_res2_bits));
    const float xh = sqrtf(__uint_as_float(*max_norm2_bits)) + xl;
    const float g = static_cast<float>(d_pad) * 2.384185791015625e-07f;   // d_pad * 2^-22
    const float err = (ql * xh + qn * xl + 2.f * g * qn * xh) * 1.0001f;
    // T = -inf: no split dropped a row and the merge kept every entry, the candidates ARE the shard
    const bool certified = t_bound == -CUDART_INF_F || (n_found >= k && kth > t_bound + err);
    need_fallback[q] = certified ? 0 : 1;
    if (!certified) {
      tile_flag[q >> 6] = 1;
      atomicAdd(n_fallback, 1);
    }
  }
}
