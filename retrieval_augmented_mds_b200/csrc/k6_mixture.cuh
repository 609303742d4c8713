// k6_mixture.cuh — generation / copy mixture, the last step of the retrieval marginalisation
// (forward): reference sotasum/retriever_generator.py:391-404
//     probs = gen_gate * softmax(logits, -1)
//     probs.scatter_add_(-1, copy_sequence (expanded over time), copy_probs)      # copy_probs = copy_gate * attention
//     outs  = log(probs + 1e-7)
// i.e. the memory documents retrieved by the search inject their tokens' copy probabilities
// (attention over the k*L memory positions, biased by the doc scores of K2) into the vocabulary
// distribution. The reference runs it as 5 full-vocabulary tensor ops (softmax, mul, scatter_add, add,
// log: ~5 reads + 4 writes of [B, T, V] fp32).
//
// One CTA PAIR (cluster of 2) per (batch, time) row, each CTA staging HALF of the vocabulary row in
// shared memory (V <= 56,000: BART / LED have 50,265): one HBM read of the logits, max and sum by block
// reductions exchanged through distributed shared memory, the S copy probabilities scattered with
// shared-memory float atomics by the CTA that owns the token (repeated tokens accumulate; their order
// is not fixed, like torch's CUDA scatter_add_), one HBM write of the log-probabilities. Half rows let
// two CTAs of DIFFERENT rows share an SM, so one row's shared-memory passes overlap the other's HBM
// traffic (a whole-row CTA per SM idled HBM during its passes: 43 % of the copy peak).
// HBM-bound: algorithmic bytes per row = 8 V + 12 S (each CTA of the pair scans the S entries: + 12 S).
#pragma once
#include "common.cuh"

#include "ptx.cuh"

namespace mix {
constexpr int THREADS = 512;
constexpr int MAX_V = 56000;                       // 2 x 28000 floats = 2 x 109.4 KiB: two CTAs per SM
constexpr int MAX_HALF = MAX_V / 2;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  __syncthreads();
  return r;
}

__device__ __forceinline__ void st_peer_f32(float* local_addr, uint32_t peer_rank, float v) {
  const uint32_t remote = ptx::mapa(ptx::smem_u32(local_addr), peer_rank);
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 2) copy_mixture_kernel(
    const float* __restrict__ logits,       // [R, V]
    const float* __restrict__ gen_gate,     // [R]
    const float* __restrict__ copy_probs,   // [R, S]
    const int64_t* __restrict__ copy_seq,   // [R / rows_per_batch, S]
    int rows_per_batch, int V, int S, float eps, float* __restrict__ out,
    float* __restrict__ stats) {            // [R, 2] row max and sum of exp(logits - max) for the backward, or null
  extern __shared__ __align__(16) float row_raw[];   // this CTA's half of the vocabulary row (+ up to 3 floats of shift)
  __shared__ float red[THREADS / 32];
  __shared__ float peer_val[2];             // written by the peer CTA: its half's max, then its half's sum
  const uint32_t rank = ptx::cluster_ctarank(), peer = rank ^ 1u;
  const int64_t r = blockIdx.x >> 1;
  const int half = (V + 1) / 2;
  const int v0 = static_cast<int>(rank) * half, n = min(V, v0 + half) - v0;   // vocabulary slice [v0, v0 + n)
  const float* src = logits + r * V + v0;
  float mx = -CUDART_INF_F;
  // Rows of an odd vocabulary (BART / LED: 50,265) are only 4-byte aligned, but every row has a 16-byte aligned
  // interior: `head` scalars, then 16-byte loads (4 in flight per thread = 32 KB per CTA: the scalar version kept
  // 16 KB in flight per CTA and sat at 46 % of the copy peak, latency bound), then < 4 tail scalars. The shared row is
  // shifted by (4 - head) % 4 floats so that the vector stores into it are aligned too.
  const int head = min(n, static_cast<int>((4u - ((reinterpret_cast<uintptr_t>(src) >> 2) & 3u)) & 3u));
  float* row = row_raw + ((4 - head) & 3);
  const int n4 = (n - head) >> 2;
  const float4* src4 = reinterpret_cast<const float4*>(src + head);
  for (int ib = threadIdx.x; ib < n4; ib += 4 * THREADS) {
    float4 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ib + i * THREADS;
      x[i] = j < n4 ? __ldg(src4 + j) : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ib + i * THREADS;
      if (j < n4) *reinterpret_cast<float4*>(row + head + 4 * j) = x[i];
      mx = fmaxf(fmaxf(mx, fmaxf(x[i].x, x[i].y)), fmaxf(x[i].z, x[i].w));
    }
  }
  {
    const int tail0 = head + 4 * n4;                 // scalars: [0, head) and [tail0, n)
    const int t = static_cast<int>(threadIdx.x);
    const int v = t < head ? t : tail0 + (t - head);
    if (v < n && (t < head || v >= tail0)) {
      const float x = __ldg(src + v);
      row[v] = x;
      mx = fmaxf(mx, x);
    }
  }
  mx = block_reduce(mx, red, true);
  if (threadIdx.x == 0) st_peer_f32(&peer_val[0], peer, mx);
  ptx::cluster_sync_all();
  mx = fmaxf(mx, peer_val[0]);
  float sum = 0.f;
  for (int v = threadIdx.x; v < n; v += THREADS) {
    const float e = __expf(row[v] - mx);
    row[v] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);
  if (threadIdx.x == 0) st_peer_f32(&peer_val[1], peer, sum);
  ptx::cluster_sync_all();
  if (stats && rank == 0 && threadIdx.x == 0) {
    stats[2 * r] = mx;
    stats[2 * r + 1] = sum + peer_val[1];
  }
  const float scale = gen_gate[r] / (sum + peer_val[1]);
  for (int v = threadIdx.x; v < n; v += THREADS) row[v] *= scale;
  __syncthreads();
  const int64_t* seq = copy_seq + (r / rows_per_batch) * S;
  const float* cp = copy_probs + r * S;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    const int64_t tok = seq[s] - v0;
    if (tok >= 0 && tok < n) atomicAdd(&row[tok], cp[s]);
  }
  __syncthreads();
  float* dst = out + r * V + v0;
  for (int v = threadIdx.x; v < n; v += THREADS) dst[v] = __logf(row[v] + eps);
}
}  // namespace mix
