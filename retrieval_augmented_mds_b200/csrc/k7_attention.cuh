// k7_attention.cuh — the score-biased copy attention's softmax and the backward of the generation / copy
// mixture: the end of the retrieval marginalisation on the TRAINING path.
//
// Reference sotasum/decoder_own.py:102-134 (LEDDecoderAttention.forward of the copy decoder, ONE head):
//     attn_weights  = bmm(query_states, key_states^T)                                   # library GEMM
//     attn_weights += beta * attention_bias.view(B, 1, -1) + beta_bias                  # attention_bias = memory_bias
//     attn_weights  = attn_weights.view(B, H, T, S) + attention_mask                    # additive, 0 / finfo.min
//     attn_weights  = softmax(attn_weights, -1)                                         # ONE softmax over all k*L tokens
// with memory_bias[b, j*L + t] = mips_scores[b, j] (retriever_generator.py:188-192): the per-document logit that
// makes the single softmax over the concatenated documents a marginalisation over documents. The reference
// runs it as 4 elementwise passes over [B*T, k*L] plus the [B, k] -> [B, k*L] expand; here it is ONE pass
// (read the GEMM's scores, write the probabilities), the document score is looked up per token
// (doc_scores[b, s / mem_len]) and the broadcast never exists. The two GEMMs around it stay library GEMMs.
//
// Backward of the softmax (dS = P * (dP - sum(P * dP))) in one pass, fused with the reduction that carries
// the gradient back to the RETRIEVER: G[b, j] = sum_t sum_{s in doc j} dS[b, t, s], from which
// d mips_scores = beta * G, d beta = sum(G * mips_scores), d beta_bias = sum(G).
//
// HBM-bound, algorithmic bytes per row: forward 8 S (+ 4 S mask, L2 resident), backward 12 S.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cattn {
constexpr int THREADS = 256;
constexpr int MAX_S = 16384;      // one row of logits lives in shared memory (64 KiB)

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

// one CTA per (batch, time) row
__global__ void __launch_bounds__(THREADS) biased_softmax_fwd_kernel(
    const float* __restrict__ scores,       // [B*T, S]  q.k (already scaled)
    const float* __restrict__ doc_scores,   // [B, n_docs] or null (no bias)
    int n_docs, int mem_len, float beta, float beta_bias,
    const float* __restrict__ beta_dev,     // device {beta, beta_bias} (parameters of a training run) or null
    const float* __restrict__ mask,         // [B, S] additive or null
    int T, int S, float* __restrict__ probs) {
  extern __shared__ float row[];
  __shared__ float red[THREADS / 32];
  const int64_t r = blockIdx.x;
  const int b = static_cast<int>(r / T);
  const float* src = scores + r * S;
  const float* ds = doc_scores ? doc_scores + static_cast<int64_t>(b) * n_docs : nullptr;
  const float* mk = mask ? mask + static_cast<int64_t>(b) * S : nullptr;
  if (beta_dev) {
    beta = __ldg(beta_dev);
    beta_bias = __ldg(beta_dev + 1);
  }
  float mx = -CUDART_INF_F;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    float x = __ldg(src + s);
    // the reference's operation order, without fused multiply-adds: (x + (beta * bias + beta_bias)) + mask
    if (ds) x = __fadd_rn(x, __fadd_rn(__fmul_rn(beta, __ldg(ds + min(s / mem_len, n_docs - 1))), beta_bias));
    if (mk) x = __fadd_rn(x, __ldg(mk + s));
    row[s] = x;
    mx = fmaxf(mx, x);
  }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    const float e = expf(row[s] - mx);
    row[s] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);
  const float inv = 1.f / sum;
  float* dst = probs + r * S;
  for (int s = threadIdx.x; s < S; s += THREADS) dst[s] = row[s] * inv;
}

__global__ void __launch_bounds__(THREADS) biased_softmax_bwd_kernel(
    const float* __restrict__ probs,        // [B*T, S]
    const float* __restrict__ dprobs,       // [B*T, S]
    int n_docs, int mem_len, int T, int S,
    float* __restrict__ dscores,            // [B*T, S]
    float* __restrict__ doc_grad) {         // [B, n_docs] += sum over the row's doc blocks of dS (zeroed by the host), or null
  extern __shared__ float bins[];           // [n_docs]
  __shared__ float red[THREADS / 32];
  const int64_t r = blockIdx.x;
  const int b = static_cast<int>(r / T);
  const float* p = probs + r * S;
  const float* dp = dprobs + r * S;
  float dot = 0.f;
  for (int s = threadIdx.x; s < S; s += THREADS) dot += __ldg(p + s) * __ldg(dp + s);
  if (doc_grad)
    for (int j = threadIdx.x; j < n_docs; j += THREADS) bins[j] = 0.f;
  dot = block_reduce(dot, red, false);      // (its barriers also publish the zeroed bins)
  float* dst = dscores + r * S;
  const int lane = threadIdx.x & 31;
  for (int s0 = threadIdx.x - lane; s0 < S; s0 += THREADS) {
    const int s = s0 + lane;
    float g = 0.f;
    if (s < S) {
      g = __ldg(p + s) * (__ldg(dp + s) - dot);
      dst[s] = g;
    }
    if (doc_grad) {
      // a warp's 32 consecutive tokens usually belong to one document: one shared-memory atomic per warp
      const int j = min(min(s, S - 1) / mem_len, n_docs - 1);
      const int j0 = __shfl_sync(0xffffffffu, j, 0), j31 = __shfl_sync(0xffffffffu, j, 31);
      if (j0 == j31) {
        float t = g;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) atomicAdd(&bins[j0], t);
      } else if (s < S) {
        atomicAdd(&bins[j], g);
      }
    }
  }
  if (doc_grad) {
    __syncthreads();
    float* dg = doc_grad + static_cast<int64_t>(b) * n_docs;
    for (int j = threadIdx.x; j < n_docs; j += THREADS)
      if (bins[j] != 0.f) atomicAdd(dg + j, bins[j]);
  }
}
}  // namespace cattn

// ---------------------------------------------------------------------------------------------------------
// Backward of the generation / copy mixture (k6_mixture.cuh; reference sotasum/retriever_generator.py:391-404):
//     out = log(m + eps),  m = gen_gate * softmax(logits) + scatter_add(copy_probs at copy_seq)
// With dM = dOut * exp(-out) (= dOut / (m + eps)) and p = softmax(logits) rebuilt from the saved row statistics:
//     d gen_gate = sum_v dM_v p_v        d logits_v = gen_gate * p_v * (dM_v - sum_u dM_u p_u)
//     d copy_probs[s] = dM[copy_seq[s]]  (0 for tokens outside [0, V))
// One CTA pair per row like the forward: each CTA keeps dM of its half of the vocabulary in shared memory
// (the gather for d copy_probs reads it there), the dot product is exchanged through DSMEM.
// HBM-bound, algorithmic bytes per row: 16 V (logits, out, dOut read; d logits written) + 12 S.
namespace mix {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 2) copy_mixture_bwd_kernel(
    const float* __restrict__ logits,       // [R, V]
    const float* __restrict__ out,          // [R, V] forward result
    const float* __restrict__ dout,         // [R, V]
    const float* __restrict__ gen_gate,     // [R]
    const float* __restrict__ stats,        // [R, 2] row max and sum of exp(logits - max) saved by the forward
    const int64_t* __restrict__ copy_seq,   // [R / rows_per_batch, S]
    int rows_per_batch, int V, int S,
    float* __restrict__ dlogits,            // [R, V]
    float* __restrict__ dgate,              // [R]
    float* __restrict__ dcopy) {            // [R, S]
  extern __shared__ __align__(16) float row_raw[];   // dM of this CTA's half of the vocabulary row (+ shift)
  __shared__ float red[THREADS / 32];
  __shared__ float peer_val[1];
  const uint32_t rank = ptx::cluster_ctarank(), peer = rank ^ 1u;
  const int64_t r = blockIdx.x >> 1;
  const int half = (V + 1) / 2;
  const int v0 = static_cast<int>(rank) * half, n = min(V, v0 + half) - v0;
  const float mx = stats[2 * r], inv_sum = 1.f / stats[2 * r + 1];
  const float g = gen_gate[r];
  const float* z = logits + r * V + v0;
  const float* o = out + r * V + v0;
  const float* d = dout + r * V + v0;
  float dot = 0.f;
  // 16-byte loads over the aligned interior of the (4-byte aligned) rows when the three arrays share the
  // alignment, which they do for contiguous tensors (k6_mixture.cuh); scalars otherwise and at the two ends
  const uintptr_t az = reinterpret_cast<uintptr_t>(z), ao = reinterpret_cast<uintptr_t>(o), ad = reinterpret_cast<uintptr_t>(d);
  const bool vec_ok = ((az ^ ao) & 15u) == 0 && ((az ^ ad) & 15u) == 0;
  const int head = vec_ok ? min(n, static_cast<int>((4u - ((az >> 2) & 3u)) & 3u)) : n;
  float* row = row_raw + ((4 - (vec_ok ? head : 0)) & 3);
  const int n4 = vec_ok ? (n - head) >> 2 : 0;
  const float4* z4 = reinterpret_cast<const float4*>(z + head);
  const float4* o4 = reinterpret_cast<const float4*>(o + head);
  const float4* d4 = reinterpret_cast<const float4*>(d + head);
  for (int ib = threadIdx.x; ib < n4; ib += 2 * THREADS) {
    float4 zz[2], oo[2], dd[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = ib + i * THREADS;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      zz[i] = j < n4 ? __ldg(z4 + j) : zero;
      oo[i] = j < n4 ? __ldg(o4 + j) : zero;
      dd[i] = j < n4 ? __ldg(d4 + j) : zero;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = ib + i * THREADS;
      if (j < n4) {
        float4 dm;
        dm.x = dd[i].x * __expf(-oo[i].x);
        dm.y = dd[i].y * __expf(-oo[i].y);
        dm.z = dd[i].z * __expf(-oo[i].z);
        dm.w = dd[i].w * __expf(-oo[i].w);
        *reinterpret_cast<float4*>(row + head + 4 * j) = dm;
        dot += dm.x * (__expf(zz[i].x - mx) * inv_sum) + dm.y * (__expf(zz[i].y - mx) * inv_sum) +
               dm.z * (__expf(zz[i].z - mx) * inv_sum) + dm.w * (__expf(zz[i].w - mx) * inv_sum);
      }
    }
  }
  {
    const int tail0 = head + 4 * n4;                 // scalars: [0, head) and [tail0, n) (everything when !vec_ok)
    for (int t = threadIdx.x; t < head + (n - tail0); t += THREADS) {
      const int v = t < head ? t : tail0 + (t - head);
      const float dm = __ldg(d + v) * __expf(-__ldg(o + v));
      row[v] = dm;
      dot += dm * (__expf(__ldg(z + v) - mx) * inv_sum);
    }
  }
  dot = block_reduce(dot, red, false);
  if (threadIdx.x == 0) st_peer_f32(&peer_val[0], peer, dot);
  ptx::cluster_sync_all();
  dot += peer_val[0];
  if (rank == 0 && threadIdx.x == 0) dgate[r] = dot;
  float* dz = dlogits + r * V + v0;
  for (int v = threadIdx.x; v < n; v += THREADS)   // logits again: the row was read a moment ago, L2 resident
    dz[v] = g * (__expf(__ldg(z + v) - mx) * inv_sum) * (row[v] - dot);
  const int64_t* seq = copy_seq + (r / rows_per_batch) * S;
  float* dc = dcopy + r * S;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    const int64_t tok = seq[s];
    const int64_t t = tok - v0;
    if (t >= 0 && t < n) dc[s] = row[t];
    else if (rank == 0 && (tok < 0 || tok >= V)) dc[s] = 0.f;   // tokens outside the vocabulary were skipped
  }
}
}  // namespace mix
