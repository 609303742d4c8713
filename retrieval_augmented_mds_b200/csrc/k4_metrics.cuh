// k4_metrics.cuh — retrieval metrics on the device.
//
// Reference: retriever_metrics (sotasum/pretrain.py:69-85, copy at retriever_lightning.py:71-87) fed
// by the hit matrix built at sotasum/mips.py:456-463: pred[b, j] = (aid of retrieved row j == aid of
// query b). The reference gathers the retrieved examples from the Arrow dataset on the host and
// builds `pred` in a Python double loop; here the ids never leave the GPU: one warp per query reads
// the row labels of its k retrieved ids and reduces recall, reciprocal rank and average precision.
//
// Semantics kept bit for bit in structure, including the reference's quirk: reciprocal_rank =
// 1 / argmax(pred) with inf -> 0, i.e. a hit at rank 1 (index 0) and a row with no hit both score 0.
#pragma once
#include "common.cuh"

// per_query[b] = {hits/counts, rr, ap}; pred_out (optional) [nq, k] receives the hit matrix.
__global__ void __launch_bounds__(128) retrieval_metrics_rows_kernel(
    const int64_t* __restrict__ ids, int nq, int k, const int64_t* __restrict__ row_aid, int64_t n_rows,
    const int64_t* __restrict__ query_aid, const float* __restrict__ counts, float* __restrict__ per_query,
    float* __restrict__ pred_out) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  const int64_t want = query_aid[q];
  // ranks in chunks of 32 (any k: multi-pass searches return more than 64 results); lane l owns rank base + l
  float n_hits = 0.f, ap = 0.f, cum = 0.f;
  int first = -1;                                  // argmax of a 0/1 row = index of the first 1, or 0 when there is none
  for (int base = 0; base < k; base += 32) {
    const int j = base + lane;
    float h = 0.f;
    if (j < k) {
      const int64_t id = ids[static_cast<size_t>(q) * k + j];
      h = (id >= 0 && id < n_rows && row_aid[id] == want) ? 1.f : 0.f;
      if (pred_out) pred_out[static_cast<size_t>(q) * k + j] = h;
    }
    uint32_t m = __ballot_sync(0xffffffffu, h != 0.f);
    n_hits += static_cast<float>(__popc(m));
    if (first < 0 && m) first = base + __ffs(m) - 1;
    // precision = cumsum(pred) / arange(1, k+1) * pred, summed
    while (m) {
      const int t = __ffs(m) - 1;
      m &= m - 1;
      cum += 1.f;
      ap += cum / static_cast<float>(base + t + 1);
    }
  }
  if (lane != 0) return;
  const float rr = first <= 0 ? 0.f : 1.f / static_cast<float>(first);   // 1/0 = inf -> masked to 0
  const float c = counts[q];
  per_query[3 * q + 0] = n_hits / c;
  per_query[3 * q + 1] = rr;
  per_query[3 * q + 2] = ap / c;
}

// out[0..2] = mean over queries, fixed reduction order (one block).
__global__ void __launch_bounds__(256) retrieval_metrics_mean_kernel(const float* __restrict__ per_query, int nq,
                                                                     float* __restrict__ out) {
  __shared__ double s[3][256];
  double a[3] = {0.0, 0.0, 0.0};
  for (int q = threadIdx.x; q < nq; q += 256)
    for (int c = 0; c < 3; ++c) a[c] += static_cast<double>(per_query[3 * q + c]);
  for (int c = 0; c < 3; ++c) s[c][threadIdx.x] = a[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int c = 0; c < 3; ++c) s[c][threadIdx.x] += s[c][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 3) out[threadIdx.x] = static_cast<float>(s[threadIdx.x][0] / static_cast<double>(nq));
}
