// k1_simt.cuh — K1 (SIMT variant): exact fp32-FMA query x bank contraction with a fused
// per-query top-k epilogue. The [nq, N] score matrix of the reference's brute force
// (sotasum/mips.py:552-560: x @ y.T then argsort) never exists; each block keeps running
// top-k lists for its 64 queries over its slice of the bank and writes [k] candidates per
// (split, sub-list, query).
//
// Used for fp32 banks (bit-faithful fp32 products and accumulation, BASELINE config 2) and as
// the bisecting reference for the tcgen05 kernel on bf16 banks.
//
// Tile: 64 queries x 128 bank rows x 16 k per step, 256 threads, 4x8 register micro-tile.
// Roofline: fp32 FMA pipe (2*nq*N*d flops); bank bytes are read once per 64-query tile.
#pragma once
#include "common.cuh"

namespace simt {
constexpr int BM = 64;    // queries per block
constexpr int BN = 128;   // bank rows per tile
constexpr int BK = 16;    // k per smem step
constexpr int SLD = 132;  // score tile leading dimension (bank-conflict-free scan)
constexpr int THREADS = 256;

__host__ __device__ inline int nsub_for_k(int k) { return k <= 16 ? 4 : (k <= 32 ? 2 : 1); }

inline size_t smem_bytes(int k) {
  const int nsub = nsub_for_k(k);
  return sizeof(float) * (2 * BK * BM + 2 * BK * BN + BM * SLD) +
         static_cast<size_t>(BM) * nsub * k * (sizeof(float) + sizeof(int));
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16);
  v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16);
  v[3] = __uint_as_float(t.y & 0xffff0000u);
}

// grid: (n_qtiles, n_splits). q has n_qtiles*64 rows (zero padded), bank rows beyond ntotal
// inside the last tile are read (capacity is a multiple of 64... the loader clamps) and masked.
template <typename T, bool kL2>
__global__ void __launch_bounds__(THREADS, 2) search_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ bank, const float* __restrict__ xnorm2, int nq,
    int64_t ntotal, int d_pad, int k, const int* __restrict__ ignore_local, int n_tiles,
    float* __restrict__ part_key, int* __restrict__ part_ids,
    const int* __restrict__ tile_active,             // [n_qtiles] or null: query tiles to (re)compute
    const float* __restrict__ after_key = nullptr,   // multi-pass search: only rows strictly after
    const int* __restrict__ after_row = nullptr) {   // (after_key, after_row) are eligible
  if (tile_active && !tile_active[blockIdx.x]) return;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);          // [2][BK][BM]
  float* Bs = As + 2 * BK * BM;                            // [2][BK][BN]
  float* S = Bs + 2 * BK * BN;                             // [BM][SLD]
  const int nsub = nsub_for_k(k);
  float* list_key = S + BM * SLD;                          // [BM*nsub][k]
  int* list_id = reinterpret_cast<int*>(list_key + BM * nsub * k);

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int qrow0 = blockIdx.x * BM;
  const int n_splits = gridDim.y, split = blockIdx.y;
  const int tile0 = static_cast<int>(static_cast<int64_t>(split) * n_tiles / n_splits);
  const int tile1 = static_cast<int>(static_cast<int64_t>(split + 1) * n_tiles / n_splits);

  // running top-k state
  const bool scanner = tid < BM * nsub;
  const int ql = tid / nsub, sub = tid - ql * nsub;
  float* lk = list_key + static_cast<size_t>(tid) * k;
  int* li = list_id + static_cast<size_t>(tid) * k;
  float thr = -CUDART_INF_F;
  int ign = -1;
  if (scanner) {
    for (int i = 0; i < k; ++i) {
      lk[i] = -CUDART_INF_F;
      li[i] = -1;
    }
    if (ignore_local && qrow0 + ql < nq) ign = ignore_local[qrow0 + ql];
  }
  float bkey = 0.f;
  int brow = -1;
  if (after_key && scanner && qrow0 + ql < nq) {
    bkey = after_key[qrow0 + ql];
    brow = after_row[qrow0 + ql];
  }

  // global->smem loader mapping: 4 consecutive k of one row per thread
  const int lrow = tid >> 2, lk4 = (tid & 3) * 4;
  const T* qa = q + static_cast<size_t>(qrow0 + lrow) * d_pad + lk4;
  const int n_ksteps = d_pad / BK;

  for (int tile = tile0; tile < tile1; ++tile) {
    const int64_t brow0 = static_cast<int64_t>(tile) * BN;
    // clamp rows so loads stay inside the shard; clamped rows are masked in the scan
    const int64_t r0 = min(brow0 + lrow, ntotal - 1), r1 = min(brow0 + lrow + 64, ntotal - 1);
    const T* xb0 = bank + r0 * d_pad + lk4;
    const T* xb1 = bank + r1 * d_pad + lk4;

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[4], rb0[4], rb1[4];
    load4<T>(qa, ra);
    load4<T>(xb0, rb0);
    load4<T>(xb1, rb1);
    for (int ks = 0; ks < n_ksteps; ++ks) {
      float* as = As + (ks & 1) * BK * BM;
      float* bs = Bs + (ks & 1) * BK * BN;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        as[(lk4 + i) * BM + lrow] = ra[i];
        bs[(lk4 + i) * BN + lrow] = rb0[i];
        bs[(lk4 + i) * BN + lrow + 64] = rb1[i];
      }
      __syncthreads();
      if (ks + 1 < n_ksteps) {
        load4<T>(qa + (ks + 1) * BK, ra);
        load4<T>(xb0 + (ks + 1) * BK, rb0);
        load4<T>(xb1 + (ks + 1) * BK, rb1);
      }
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(as + kk * BM + ty * 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * BN + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(bs + kk * BN + 64 + tx * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      // double-buffered smem: the next step writes the other buffer, one barrier per step
    }

    // scores -> smem tile (ranking key: <q,x> or <q,x> - |x|^2/2)
    float adj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      adj[j] = 0.f;
      if (kL2) {
        const int64_t r = brow0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        adj[j] = r < ntotal ? -0.5f * xnorm2[r] : 0.f;
      }
    }
    __syncthreads();  // previous tile's scan is done with S
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* srow = S + (ty * 4 + i) * SLD;
      *reinterpret_cast<float4*>(srow + tx * 4) =
          make_float4(acc[i][0] + adj[0], acc[i][1] + adj[1], acc[i][2] + adj[2], acc[i][3] + adj[3]);
      *reinterpret_cast<float4*>(srow + 64 + tx * 4) =
          make_float4(acc[i][4] + adj[4], acc[i][5] + adj[5], acc[i][6] + adj[6], acc[i][7] + adj[7]);
    }
    __syncthreads();

    if (scanner && qrow0 + ql < nq) {
      const float* srow = S + ql * SLD;
      for (int c = sub; c < BN; c += nsub) {
        const float s = srow[c];
        if (s > thr) {
          const int64_t id = brow0 + c;
          const bool eligible = !after_key || s < bkey || (s == bkey && id > brow);
          if (id < ntotal && id != ign && eligible) thr = topk_list_insert(lk, li, k, s, static_cast<int>(id));
        }
      }
    }
  }

  if (scanner && qrow0 + ql < nq) {
    const size_t p = static_cast<size_t>(split) * nsub + sub;
    const size_t o = (p * nq + (qrow0 + ql)) * k;
    for (int i = 0; i < k; ++i) {
      part_key[o + i] = lk[i];
      part_ids[o + i] = li[i];
    }
  }
}
}  // namespace simt
