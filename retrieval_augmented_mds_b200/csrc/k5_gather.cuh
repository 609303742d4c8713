// k5_gather.cuh — result gather for the consumer of the search.
//
// After Mips.search the reference looks the retrieved rows up in the Arrow dataset on the host,
// re-tokenises their TEXT every step (sotasum/mips.py:428, :465-505: memory_tokenizer(flat_texts,
// padding="max_length", truncation=True)), derives memory_attention_mask (attention_mask with the
// <s> / </s> positions cleared, :494-501) and global_attention_mask (1 on the first token,
// :483-487). With the memory tokenised ONCE into an HBM store [N, L] these four tensors are a gather
// by the ids that are already on the GPU. gather_rows does the same for the bank rows themselves
// ([n, d] fp32), so that doc scores can be recomputed with gradient w.r.t. the query when the
// memory encoder is frozen (retriever_generator.py:158-172).
//
// HBM-bound copies: bytes = n * L * (4 read + 32 written) resp. n * d * (e + 4).
#pragma once
#include "common.cuh"

// One block per gathered row. ids < 0 (k > ntotal padding) or outside the store give a row of
// pad tokens with all masks 0.
__global__ void __launch_bounds__(128) gather_tokens_kernel(
    const int32_t* __restrict__ store_ids, const int32_t* __restrict__ store_len, int64_t n_rows, int L,
    const int64_t* __restrict__ ids, int32_t pad_id, int32_t bos_id, int32_t eos_id,
    int64_t* __restrict__ input_ids, int64_t* __restrict__ attention_mask,
    int64_t* __restrict__ memory_attention_mask, int64_t* __restrict__ global_attention_mask) {
  const int64_t r = blockIdx.x;
  const int64_t id = ids[r];
  const bool ok = id >= 0 && id < n_rows;
  const int32_t* src = store_ids + (ok ? id : 0) * L;
  const int len = ok ? store_len[id] : 0;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const int32_t tok = ok ? src[t] : pad_id;
    const int64_t att = t < len ? 1 : 0;
    const size_t o = static_cast<size_t>(r) * L + t;
    input_ids[o] = tok;
    attention_mask[o] = att;
    if (memory_attention_mask) memory_attention_mask[o] = (tok == bos_id || tok == eos_id) ? 0 : att;
    if (global_attention_mask) global_attention_mask[o] = (t == 0) ? 1 : 0;
  }
}

// One warp per gathered row: stored row (bf16 or fp32) -> fp32 [d]; rows outside this shard
// ([id_offset, id_offset + ntotal)) are written as zeros (a row-sharded bank sums the ranks' outputs).
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ bank, int64_t ntotal, int d,
                                                          int d_pad, const int64_t* __restrict__ ids, int64_t n,
                                                          int64_t id_offset, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t row = ids[r] - id_offset;
  const bool ok = ids[r] >= 0 && row >= 0 && row < ntotal;
  const T* src = bank + (ok ? row : 0) * d_pad;
  float* dst = out + r * d;
  for (int c = lane; c < d; c += 32) dst[c] = ok ? to_f32<T>(src[c]) : 0.f;
}
