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
  extern __shared__ float row[];            // this CTA's half of the vocabulary row
  __shared__ float red[THREADS / 32];
  __shared__ float peer_val[2];             // written by the peer CTA: its half's max, then its half's sum
  const uint32_t rank = ptx::cluster_ctarank(), peer = rank ^ 1u;
  const int64_t r = blockIdx.x >> 1;
  const int half = (V + 1) / 2;
  const int v0 = static_cast<int>(rank) * half, n = min(V, v0 + half) - v0;   // vocabulary slice [v0, v0 + n)
  const float* src = logits + r * V + v0;
  float mx = -CUDART_INF_F;
  // 8 independent loads in flight per thread (a rolled loop keeps ~1: 4 KB in flight per SM starves HBM)
  for (int vb = threadIdx.x; vb < n; vb += 8 * THREADS) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = vb + i * THREADS;
      x[i] = v < n ? __ldg(src + v) : -CUDART_INF_F;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = vb + i * THREADS;
      if (v < n) row[v] = x[i];
      mx = fmaxf(mx, x[i]);
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
