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
// One CTA per (batch, time) row with the WHOLE vocabulary row staged in shared memory (V <= 57,000:
// BART / LED have 50,265): one HBM read of the logits, max and sum by block reductions, the S copy
// probabilities scattered with shared-memory float atomics (repeated tokens accumulate; their
// order is not fixed, like torch's CUDA scatter_add_), one HBM write of the log-probabilities.
// HBM-bound: algorithmic bytes per row = 8 V + 12 S.
#pragma once
#include "common.cuh"

namespace mix {
constexpr int THREADS = 512;
constexpr int MAX_V = 57000;   // 57000 * 4 B = 222.7 KiB of the 227 KiB opt-in shared memory

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

__global__ void __launch_bounds__(THREADS, 1) copy_mixture_kernel(
    const float* __restrict__ logits,       // [R, V]
    const float* __restrict__ gen_gate,     // [R]
    const float* __restrict__ copy_probs,   // [R, S]
    const int64_t* __restrict__ copy_seq,   // [R / rows_per_batch, S]
    int rows_per_batch, int V, int S, float eps, float* __restrict__ out) {
  extern __shared__ float row[];            // [V]
  __shared__ float red[THREADS / 32];
  const int64_t r = blockIdx.x;
  const float* src = logits + r * V;
  float mx = -CUDART_INF_F;
  for (int v = threadIdx.x; v < V; v += THREADS) {
    const float x = src[v];
    row[v] = x;
    mx = fmaxf(mx, x);
  }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int v = threadIdx.x; v < V; v += THREADS) {
    const float e = __expf(row[v] - mx);
    row[v] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);
  const float scale = gen_gate[r] / sum;
  for (int v = threadIdx.x; v < V; v += THREADS) row[v] *= scale;
  __syncthreads();
  const int64_t* seq = copy_seq + (r / rows_per_batch) * S;
  const float* cp = copy_probs + r * S;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    const int64_t tok = seq[s];
    if (tok >= 0 && tok < V) atomicAdd(&row[tok], cp[s]);
  }
  __syncthreads();
  float* dst = out + r * V;
  for (int v = threadIdx.x; v < V; v += THREADS) dst[v] = __logf(row[v] + eps);
}
}  // namespace mix
