// k2_merge.cuh — K2: k-way merge of per-split / per-rank candidate lists, one warp per query.
//
// Selection is "k_out rounds of warp arg-max under a strictly-after constraint": round j picks
// the best candidate that comes strictly after round j-1's pick in the total order
// (key descending, id ascending). No scratch, no mutation of the inputs, deterministic for
// any number of parts (shard-count invariant), and the ignore filter of Mips.search
// (reference sotasum/mips.py:388-398: drop the hit whose id == ignore_indexes[j]) is a skip.
//
// LOCAL = true : candidates carry int32 shard-local rows; output is still the ranking key,
//                ids become global (id_offset + row) and |x|^2 is gathered from the shard.
// LOCAL = false: candidates carry int64 global ids (+ optional |x|^2); output goes through the
//                metric transform and the doc-score arithmetic of
//                retriever_generator.py:158-193 (cosine, per-doc softmax, memory_bias).
#pragma once
#include "common.cuh"
#include "mips_b200.h"

// One candidate as it travels through the all-gather: 16 bytes.
struct __align__(16) PackedCand {
  float key;
  float xn2;
  int64_t id;
};

// Peer-memory exchange fused into the merge kernels (one box, NVLink/NVSwitch, CUDA IPC mappings):
// the LOCAL merge of rank r stores its [nq, k] records straight into slot r of EVERY rank's exchange
// buffer and the last warp to finish raises rank r's arrival flag on every rank; the FINAL merge of
// each rank waits for the G flags of this search (`seq`) and merges what landed in its own buffer.
// No collective launch, no host involvement; the only traffic is 16 * nq * k bytes per peer.
struct XchgOut {
  PackedCand* const* bufs;   // device array [n_peers]: this rank's region on every rank (itself included)
  uint32_t* const* flags;    // device array [n_peers]: this rank's arrival flag on every rank
  unsigned int* done;        // local counter of finished queries (returns to 0 at the end of the launch)
  uint32_t seq;
  int n_peers;               // 0: no exchange
};
struct XchgIn {
  const uint32_t* flags;     // [n_ranks] arrival flags in THIS rank's buffer
  uint32_t seq;
  int n_ranks;               // 0: no exchange
  uint32_t* timeout_flag;    // receives `seq` when a peer's list did not arrive in time (host polls it)
  unsigned long long timeout_ns;
};
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct MergeBest {
  float key;
  int64_t id;
  int pos;
};

__device__ __forceinline__ bool merge_better(float ka, int64_t ia, float kb, int64_t ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

template <bool LOCAL>
__global__ void __launch_bounds__(128) merge_topk_kernel(
    const float* __restrict__ cand_key, const void* __restrict__ cand_ids_v,
    const float* __restrict__ cand_xn2,   // !LOCAL: [n_parts, nq, k_in] or null
    const float* __restrict__ bank_xn2,   // LOCAL: shard norm array
    int n_parts, int nq, int k_in, int k_out, int64_t id_offset,
    const int64_t* __restrict__ ignore_ids, int metric, int out_mode, float phi,
    const float* __restrict__ q_norm2, float* __restrict__ out_key, int64_t* __restrict__ out_ids,
    float* __restrict__ out_xn2, float* __restrict__ cosine, float* __restrict__ doc_prob,
    float beta, float beta_bias, float* __restrict__ memory_bias, int mem_len,
    const PackedCand* __restrict__ cand_packed,   // !LOCAL: packed input instead of the 3 arrays
    PackedCand* __restrict__ out_packed,          // LOCAL: packed output instead of the 3 arrays
    const int* __restrict__ q_active = nullptr,   // only these queries are merged (others keep their outputs)
    int stage_cap = 0,                            // shared-memory staging entries per warp (dynamic smem: 8 bytes per
                                                  // entry LOCAL, 16 bytes FINAL); 0 = read candidates from global memory
    XchgOut xo = XchgOut{nullptr, nullptr, nullptr, 0u, 0},   // LOCAL: publish to the peers
    XchgIn xi = XchgIn{nullptr, 0u, 0, nullptr, 0ull},       // FINAL: wait for the peers
    int cut_above = 512) {                                    // LOCAL: radix-cut the staged set when it is larger
  // 2 * MIPS_MAX_K: the candidate merge of the exact fp32 search keeps up to 128 entries per query
  __shared__ float s_key[4][2 * MIPS_MAX_K];
  __shared__ float s_xn2[4][2 * MIPS_MAX_K];
  __shared__ float s_cos[4][2 * MIPS_MAX_K];
  __shared__ int64_t s_id[4][2 * MIPS_MAX_K];
  __shared__ int s_pos[4][2 * MIPS_MAX_K];
  __shared__ unsigned int s_hist[4][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + w;   // 4 queries per block, fewer when the staging area is large
  if (q >= nq) return;
  if (q_active && !q_active[q]) return;
  if (!LOCAL && xi.n_ranks > 0) {
    // every rank's list of this search has landed in this rank's buffer. The wait is bounded in wall time
    // (ordinary rank skew — a checkpoint, a data-loader stall — is minutes at worst); when it expires the
    // context stays healthy: the query returns "no result" (ids -1) and the time-out flag tells the host,
    // which falls back to the NCCL exchange (ShardedFlatIndex.check_exchange)
    bool lost = false;
    if (lane < xi.n_ranks) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(xi.flags + lane) != xi.seq) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > xi.timeout_ns) {
          lost = true;
          break;
        }
      }
    }
    lost = __any_sync(0xffffffffu, lost);
    if (lost) {
      if (lane == 0 && xi.timeout_flag) atomicExch(xi.timeout_flag, xi.seq);
      for (int j = lane; j < k_out; j += 32) {
        const size_t o = static_cast<size_t>(q) * k_out + j;
        out_ids[o] = -1;
        out_key[o] = (out_mode == MIPS_OUT_IP) ? -CUDART_INF_F : CUDART_INF_F;
        if (cosine) cosine[o] = 0.f;
        if (doc_prob) doc_prob[o] = 0.f;
      }
      if (memory_bias)
        for (int t = lane; t < k_out * mem_len; t += 32) memory_bias[static_cast<size_t>(q) * k_out * mem_len + t] = 0.f;
      return;
    }
  }

  const int32_t* ids32 = static_cast<const int32_t*>(cand_ids_v);
  const int64_t* ids64 = static_cast<const int64_t*>(cand_ids_v);
  const int64_t ign = ignore_ids ? ignore_ids[q] : -1;
  const int C = n_parts * k_in;

  // Every candidate of the query is read from global memory ONCE, into shared memory (the launch provides
  // stage_cap entries per warp): LOCAL as {key bits, shard-local row}, FINAL as {key bits, position} + the int64
  // id. The first version re-read them in every selection round — k_out dependent L2 round trips per query
  // (ncu launch lists: 28 us for the final merge of 1024 x 32 results, 45 us for a local merge).
  extern __shared__ __align__(16) unsigned char s_dyn[];
  uint2* stage = reinterpret_cast<uint2*>(s_dyn) + static_cast<size_t>(w) * stage_cap;
  int64_t* stage_id = nullptr;          // FINAL only
  if (!LOCAL) stage_id = reinterpret_cast<int64_t*>(s_dyn + static_cast<size_t>(blockDim.x >> 5) * stage_cap * sizeof(uint2)) +
                         static_cast<size_t>(w) * stage_cap;
  const bool staged = stage_cap >= C && C > 0;
  if (staged) {
    for (int c = lane; c < C; c += 32) {
      const int p = c / k_in, sidx = c - p * k_in;
      const size_t a = (static_cast<size_t>(p) * nq + q) * k_in + sidx;
      if (LOCAL) {
        stage[c] = make_uint2(__float_as_uint(cand_key[a]), static_cast<uint32_t>(ids32[a]));
      } else if (cand_packed) {
        const PackedCand pc = cand_packed[a];
        stage[c] = make_uint2(__float_as_uint(pc.key), static_cast<uint32_t>(c));
        stage_id[c] = pc.id;
      } else {
        stage[c] = make_uint2(__float_as_uint(cand_key[a]), static_cast<uint32_t>(c));
        stage_id[c] = ids64[a];
      }
    }
    __syncwarp();
  }
  int C_eff = C;
  if (LOCAL && staged && C > cut_above) {
    // Cut the staged candidates down to the k_out best (plus ties at the cut) BEFORE the ordered
    // selection rounds (148 splits x 64 entries x 64 rounds took 3.7 ms for 128 queries): a 4-pass, 8-bit MSB
    // radix select on the order-preserving integer image of the keys finds the k_out-th largest key T in
    // O(candidates), then one pass compacts the entries with key >= T to the front of the staging area.
    // Exact: ties at T all survive and the rounds below order them by id.
    unsigned int* hist = s_hist[w];
    uint32_t prefix = 0u, known = 0u;
    int remaining = k_out;
    bool cut = true;
    for (int pass = 0; pass < 4 && cut; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = lane; i < 256; i += 32) hist[i] = 0u;
      __syncwarp();
      for (int c = lane; c < C; c += 32) {
        const uint2 e = stage[c];
        if (static_cast<int32_t>(e.y) < 0) continue;
        const uint32_t u = f32_to_ordered(__uint_as_float(e.x));
        if ((u & known) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
      }
      __syncwarp();
      // lane l owns bins [8l, 8l + 8); walk from the top bin down until `remaining` entries are covered
      unsigned int mine[8], sum = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mine[i] = hist[8 * lane + i];
        sum += mine[i];
      }
      // suffix sum over lanes: incl_l = sum_{l' >= l} sum_l', above_l = entries in the bins of higher lanes
      unsigned int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += t;
      }
      const unsigned int above = incl - sum;
      const unsigned int total = __shfl_sync(0xffffffffu, incl, 0);
      if (pass == 0 && total <= static_cast<unsigned int>(k_out)) {
        cut = false;   // not more valid candidates than outputs: nothing to cut
        break;
      }
      const bool here = above < static_cast<unsigned int>(remaining) && static_cast<unsigned int>(remaining) <= above + sum;
      const uint32_t owner_mask = __ballot_sync(0xffffffffu, here);
      const int owner = __ffs(owner_mask) - 1;
      int bin = 0;
      unsigned int gt = 0u;
      if (lane == owner) {
        unsigned int acc = above;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
          if (acc < static_cast<unsigned int>(remaining) && static_cast<unsigned int>(remaining) <= acc + mine[i]) {
            bin = 8 * lane + i;
            gt = acc;
          }
          acc += mine[i];
        }
      }
      bin = __shfl_sync(0xffffffffu, bin, owner);
      gt = __shfl_sync(0xffffffffu, gt, owner);
      remaining -= static_cast<int>(gt);
      prefix |= static_cast<uint32_t>(bin) << shift;
      known |= 255u << shift;
      __syncwarp();
    }
    if (cut) {   // prefix is the ordered image of T
      int w_pos = 0;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        uint2 e = make_uint2(0u, 0xffffffffu);
        if (c < C) e = stage[c];
        const bool keep = c < C && static_cast<int32_t>(e.y) >= 0 && f32_to_ordered(__uint_as_float(e.x)) >= prefix;
        const uint32_t kept = __ballot_sync(0xffffffffu, keep);
        if (keep) stage[w_pos + __popc(kept & ((1u << lane) - 1u))] = e;   // w_pos <= c0: in place is safe
        w_pos += __popc(kept);
      }
      __syncwarp();
      C_eff = w_pos;
    }
  }

  // Selection: k_out rounds; round j picks the best candidate strictly after round j-1's pick in the total order
  // (key descending, id ascending) — no mutation, duplicates of one (key, id) collapse, shard-count invariant.
  // Every lane caches the best of ITS candidates that is still eligible; a round is two or three warp REDUX
  // instructions (max of the ordered keys, min of the ids among the lanes that hold that key) and only the lanes
  // whose cached candidate was just picked rescan their C / 32 entries.
  auto scan = [&](uint32_t pk, int64_t pid, uint32_t& bk, int64_t& bid, int& bpos) {
    bk = 0u;
    bid = INT64_MAX;
    bpos = -1;
    for (int c = lane; c < C_eff; c += 32) {
      float key;
      int64_t id;
      int pos;
      if (staged) {
        const uint2 e = stage[c];
        key = __uint_as_float(e.x);
        if (LOCAL) {
          const int32_t l = static_cast<int32_t>(e.y);
          id = l < 0 ? -1 : id_offset + l;
          pos = c;
        } else {
          id = stage_id[c];
          pos = static_cast<int>(e.y);
        }
      } else {
        const int p = c / k_in, sidx = c - p * k_in;
        const size_t a = (static_cast<size_t>(p) * nq + q) * k_in + sidx;
        pos = c;
        if (LOCAL) {
          const int32_t l = ids32[a];
          id = l < 0 ? -1 : id_offset + l;
          key = cand_key[a];
        } else if (cand_packed) {
          const PackedCand pc = cand_packed[a];
          id = pc.id;
          key = pc.key;
        } else {
          id = ids64[a];
          key = cand_key[a];
        }
      }
      if (id < 0 || id == ign) continue;
      const uint32_t ku = f32_to_ordered(key);
      if (ku == 0u) continue;                                   // (no finite or infinite float maps to 0)
      const bool after = ku < pk || (ku == pk && id > pid);
      if (!after) continue;
      if (ku > bk || (ku == bk && id < bid)) {
        bk = ku;
        bid = id;
        bpos = pos;
      }
    }
  };
  uint32_t bk;
  int64_t bid;
  int bpos;
  scan(0xffffffffu, -1, bk, bid, bpos);
  int n_found = 0;
  for (int j = 0; j < k_out; ++j) {
    const uint32_t wk = __reduce_max_sync(0xffffffffu, bk);
    if (wk == 0u) break;
    const bool has = bk == wk;
    const uint32_t whi = __reduce_min_sync(0xffffffffu, has ? static_cast<uint32_t>(static_cast<uint64_t>(bid) >> 32) : 0xffffffffu);
    const bool has2 = has && static_cast<uint32_t>(static_cast<uint64_t>(bid) >> 32) == whi;
    const uint32_t wlo = __reduce_min_sync(0xffffffffu, has2 ? static_cast<uint32_t>(static_cast<uint64_t>(bid)) : 0xffffffffu);
    const int64_t wid = static_cast<int64_t>((static_cast<uint64_t>(whi) << 32) | wlo);
    const bool mine = has2 && static_cast<uint32_t>(static_cast<uint64_t>(bid)) == wlo;
    const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    const int wpos = __shfl_sync(0xffffffffu, bpos, src);
    if (lane == 0) {
      s_key[w][j] = ordered_to_f32(wk);
      s_id[w][j] = wid;
      s_pos[w][j] = wpos;
    }
    n_found = j + 1;
    if (mine) scan(wk, wid, bk, bid, bpos);                     // (duplicates of the pick rescan too)
  }
  __syncwarp();

  // per-pick outputs, in parallel over the picks: |x|^2 gather, ids, packed records / peer buffers
  for (int j = lane; j < n_found; j += 32) {
    const int64_t id = s_id[w][j];
    const int pos = s_pos[w][j];
    float xn = 0.f;
    if (LOCAL) {
      if (bank_xn2) xn = bank_xn2[id - id_offset];
    } else {
      const int p = pos / k_in, sidx = pos - p * k_in;
      const size_t a = (static_cast<size_t>(p) * nq + q) * k_in + sidx;
      if (cand_packed) xn = cand_packed[a].xn2;
      else if (cand_xn2) xn = cand_xn2[a];
    }
    s_xn2[w][j] = xn;
    const size_t o = static_cast<size_t>(q) * k_out + j;
    if (LOCAL && xo.n_peers > 0) {
      const PackedCand rec{s_key[w][j], xn, id};
      for (int g = 0; g < xo.n_peers; ++g) xo.bufs[g][o] = rec;
    } else if (LOCAL && out_packed) {
      out_packed[o] = PackedCand{s_key[w][j], xn, id};
    } else {
      out_ids[o] = id;
    }
  }
  __syncwarp();

  const float qn2 = q_norm2 ? q_norm2[q] : 0.f;
  float lmax = -CUDART_INF_F;
  for (int j = lane; j < k_out; j += 32) {
    const size_t o = static_cast<size_t>(q) * k_out + j;
    if (j >= n_found) {
      if (LOCAL && xo.n_peers > 0) {
        for (int g = 0; g < xo.n_peers; ++g) xo.bufs[g][o] = PackedCand{-CUDART_INF_F, 0.f, -1};
        continue;
      }
      if (LOCAL && out_packed) {
        out_packed[o] = PackedCand{-CUDART_INF_F, 0.f, -1};
        continue;
      }
      out_ids[o] = -1;
      if (LOCAL) {
        out_key[o] = -CUDART_INF_F;
        if (out_xn2) out_xn2[o] = 0.f;
      } else {
        out_key[o] = (out_mode == MIPS_OUT_IP) ? -CUDART_INF_F : CUDART_INF_F;
        if (cosine) cosine[o] = 0.f;
      }
      s_cos[w][j] = -CUDART_INF_F;
      continue;
    }
    const float key = s_key[w][j], xn = s_xn2[w][j];
    if (LOCAL) {
      if (!out_packed && xo.n_peers == 0) {
        out_key[o] = key;
        if (out_xn2) out_xn2[o] = xn;
      }
      continue;
    }
    // ranking key -> inner product
    const float ip = (metric == MIPS_METRIC_L2) ? key + 0.5f * xn : key;
    float dval;
    if (out_mode == MIPS_OUT_IP) {
      dval = ip;
    } else if (out_mode == MIPS_OUT_L2) {
      dval = fmaxf(qn2 + xn - 2.f * ip, 0.f);
    } else {
      dval = qn2 + phi - 2.f * ip;
    }
    out_key[o] = dval;
    // retriever_generator.py:159-172: <q,d> / (|q| |d|)
    const float den = sqrtf(qn2) * sqrtf(xn);
    const float cs = den > 0.f ? ip / den : 0.f;
    if (cosine) cosine[o] = cs;
    s_cos[w][j] = cs;
    lmax = fmaxf(lmax, beta * cs + beta_bias);
  }
  if (LOCAL) {
    if (xo.n_peers > 0) {
      // this query's records are on their way to every peer; the last query to get here raises the flags
      __syncwarp();
      __threadfence_system();
      if (lane == 0) {
        const unsigned int ticket = atomicAdd(xo.done, 1u);
        if (ticket == static_cast<unsigned int>(nq) - 1u) {
          __threadfence_system();
          *xo.done = 0u;
          for (int g = 0; g < xo.n_peers; ++g) st_release_sys(xo.flags[g], xo.seq);
        }
      }
    }
    return;
  }
  __syncwarp();

  if (doc_prob) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    float lsum = 0.f;
    for (int j = lane; j < n_found; j += 32) lsum += __expf(beta * s_cos[w][j] + beta_bias - lmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    for (int j = lane; j < k_out; j += 32) {
      const size_t o = static_cast<size_t>(q) * k_out + j;
      doc_prob[o] = j < n_found ? __expf(beta * s_cos[w][j] + beta_bias - lmax) / lsum : 0.f;
    }
  }
  if (memory_bias) {
    // retriever_generator.py:188-192: bias[q, j*L + t] = cosine[q, j]
    const int total = k_out * mem_len;
    float* dst = memory_bias + static_cast<size_t>(q) * total;
    for (int t = lane; t < total; t += 32) {
      const int j = t / mem_len;
      dst[t] = j < n_found ? s_cos[w][j] : 0.f;
    }
  }
}
