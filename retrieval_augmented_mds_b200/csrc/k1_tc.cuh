// k1_tc.cuh — K1 (tensor-core variant): sm_100a tcgen05/TMEM/TMA query x bank contraction with a
// fused per-query top-k epilogue. Replaces the arithmetic of faiss.IndexFlat*.search as reached
// from Mips.search (reference sotasum/mips.py:383-386) and of inner_product (mips.py:552-560).
//
// Mapping (one CTA per SM, persistent over its slice of the bank):
//   * 128 queries per CTA = the 128 TMEM lanes. The query tile is STATIONARY: it is written once
//     into TMEM columns [128, 128 + d_pad/2) as packed bf16 pairs and used as the A operand
//     (tcgen05.mma with A in TMEM), so the only streamed operand is the bank.
//   * bank rows stream HBM/L2 -> shared memory by TMA (ACC_N rows x 64 k boxes, 128-byte
//     swizzle), SKCH boxes per pipeline stage, mbarrier full/empty ring.
//   * scores accumulate in fp32 in the 128 TMEM columns left beside the query tile. Two layouts
//     are compiled (template ACC_N); measured on B200 (scripts/mma_rate.cu, scripts/ldtm_rate.cu):
//     a tcgen05.mma with A in TMEM costs N/2 + ~11 cycles whatever the accumulator chaining
//     (N=64: 43.5 cycles = 73.5 % of nominal, N=128: 74 = 86.5 %, N=256: 138 = 92.7 %).
//       ACC_N = 64 (default): two 64-row accumulators, the MMA of tile t+1 overlaps the epilogue of
//                    tile t. Ceiling 73.5 % of the nominal tensor rate per clock.
//       ACC_N = 128: one 128-row accumulator; faster instructions but the epilogue drain and both
//                    barrier hand-offs are exposed once per tile (measured slower end to end:
//                    56.0k vs 68.9k queries/s on the 10M x 768 workload).
//     A wider double-buffered accumulator does not fit: the stationary 128 x 768 bf16 query tile
//     alone needs 384 of the 512 TMEM columns.
//   * epilogue: thread <-> query. Each thread pulls its lane's ACC_N scores (tcgen05.ld 32x32b),
//     releases the accumulator, takes one max over them and compares with its running k-th best;
//     only when something beats it (rare after warm-up) does it admit candidates into its unsorted
//     top-k set in shared memory (k1_topk.cuh). The [nq, N] score matrix never reaches HBM.
//
// Grid = n_qtiles * n_splits CTAs (<= #SMs): CTA (qtile, split) scans bank tiles
// [split*T/S, (split+1)*T/S) for query tile qtile; CTAs of one split run side by side so a bank
// tile is fetched from HBM once and served from L2 to the other query tiles.
//
// Roofline: tensor bound, 2*128*ACC_N*d_pad flops per tile; HBM bytes = one pass over
// the bank per batch of <= 128*n_qtiles queries.
#pragma once
#include "common.cuh"
#include "k1_topk.cuh"
#include "ptx.cuh"

namespace tc {
constexpr int BLOCK_M = 128;                          // queries per CTA (TMEM lanes)
constexpr int ACC_COLS = 128;                         // TMEM columns available for accumulators
constexpr int KCH = 64;                               // bf16 per 128-byte swizzle row
constexpr int TMEM_COLS = 512;
constexpr int Q_COL0 = ACC_COLS;                      // query tile starts after the accumulators
constexpr int MAX_DPAD = (TMEM_COLS - Q_COL0) * 2;    // 768
constexpr int THREADS = 192;                          // 4 epilogue warps + TMA warp + MMA warp
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;                    // 227 KiB opt-in maximum

// shared memory: [align pad 1024][stages][lists][barriers]
__host__ __device__ constexpr int box_bytes(int acc_n) { return acc_n * KCH * 2; }
__host__ __device__ inline int list_bytes(int k) { return BLOCK_M * topk_kcap(k) * 8; }
__host__ __device__ inline int bar_bytes() { return (2 * MAX_STAGES + 6) * 8; }
// boxes (k chunks of 64) per stage: ~48 KiB stages when that divides d_pad/64, else ~32 KiB
inline int pick_skch(int d_pad, int acc_n) {
  const int n_kch = d_pad / KCH;
  if (acc_n == 128) return n_kch % 3 == 0 ? 3 : 2;
  return n_kch % 6 == 0 ? 6 : 4;
}
inline int pick_stages(int k, int skch, int acc_n) {
  int s = (SMEM_LIMIT - 1024 - list_bytes(k) - bar_bytes()) / (skch * box_bytes(acc_n));
  return s > MAX_STAGES ? MAX_STAGES : s;
}
inline size_t smem_bytes(int k, int stages, int skch, int acc_n) {
  return 1024 + static_cast<size_t>(stages) * skch * box_bytes(acc_n) + list_bytes(k) + bar_bytes();
}

struct Params {
  const __nv_bfloat16* q;   // [n_qtiles*128, d_pad] prepared queries (zero padded)
  const float* xnorm2;      // [capacity] (L2 only)
  const int* ignore_local;  // [nq] or null
  const float* after_key;   // [nq] or null: multi-pass search, only rows strictly after (after_key, after_row) ...
  const int* after_row;     // ... in the order (key descending, row ascending) are eligible
  float* part_key;          // [n_splits, nq, k]
  int* part_ids;
  int64_t ntotal;
  int nq, d_pad, k, n_tiles, n_qtiles, n_splits, stages;
  unsigned long long cache_hint;
};

template <bool kL2, int ACC_N, int SKCH>
__global__ void __launch_bounds__(THREADS, 1) search_tc_kernel(
    const __grid_constant__ CUtensorMap tmap, const Params p) {
  static_assert(ACC_N == 64 || ACC_N == 128, "accumulator width");
  constexpr int NACC = ACC_COLS / ACC_N;                // 2 (double buffered) or 1
  constexpr int BOX_BYTES = box_bytes(ACC_N);
  constexpr int STAGE_KCH = SKCH;
  constexpr int STAGE_BYTES = SKCH * BOX_BYTES;
  constexpr int NGRP = ACC_N / 32;                      // tcgen05.ld.x32 groups per tile

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;   // 128B swizzle atoms need 1024B alignment
  uint8_t* gen = smem_raw + (base - raw_addr);

  const int S = p.stages;
  const uint32_t lists_off = static_cast<uint32_t>(S) * STAGE_BYTES;
  uint32_t* lists = reinterpret_cast<uint32_t*>(gen + lists_off);   // per warp: keys [kcap][32], ids [kcap][32]
  const uint32_t bars = base + lists_off + list_bytes(p.k);
  auto full_bar = [&](int i) { return bars + 8u * i; };
  auto empty_bar = [&](int i) { return bars + 8u * (MAX_STAGES + i); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t qready_bar = bars + 8u * (2 * MAX_STAGES + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 5);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + lists_off + list_bytes(p.k) + 8 * (2 * MAX_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x % p.n_qtiles, split = blockIdx.x / p.n_qtiles;
  const int tile0 = static_cast<int>(static_cast<int64_t>(split) * p.n_tiles / p.n_splits);
  const int tile1 = static_cast<int>(static_cast<int64_t>(split + 1) * p.n_tiles / p.n_splits);
  const int n_kch = p.d_pad / KCH;
  const int n_kstages = (n_kch + STAGE_KCH - 1) / STAGE_KCH;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tensormap(&tmap);
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(full_bar(i), 1);    // producer's arrive.expect_tx
      ptx::mbar_init(empty_bar(i), 1);   // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);    // tcgen05.commit
      ptx::mbar_init(tempty_bar(a), 4);   // one elected lane per epilogue warp
    }
    ptx::mbar_init(qready_bar, BLOCK_M);
    ptx::fence_mbar_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // accumulator a of iteration `it`: buffer index and the parity of its full/empty barriers
  auto acc_of = [](int it) { return NACC == 2 ? (it & 1) : 0; };
  auto acc_par = [](int it) { return static_cast<uint32_t>(NACC == 2 ? ((it >> 1) & 1) : (it & 1)); };

  // The producer and MMA warps run their loops with all 32 lanes converged and elect one lane only
  // around the asynchronous instructions: descriptors and addresses then live in uniform
  // registers and each tcgen05.mma / TMA costs a handful of issue slots instead of a per-lane
  // "waterfall" loop (measured: 80 cycles per MMA when the loop ran under `if (lane == 0)`).
  if (warp == 4) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      const int row0 = tile * ACC_N;
      for (int ks = 0; ks < n_kstages; ++ks) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        const int nk = min(STAGE_KCH, n_kch - ks * STAGE_KCH);
        const uint32_t dst = base + static_cast<uint32_t>(stage) * STAGE_BYTES;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(nk) * BOX_BYTES);
#pragma unroll
          for (int c = 0; c < STAGE_KCH; ++c)
            if (c < nk)
              ptx::tma_load_2d_hint(dst + c * BOX_BYTES, &tmap, full_bar(stage),
                                    (ks * STAGE_KCH + c) * KCH, row0, p.cache_hint);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    // The barrier of the NEXT stage (and of the next tile's accumulator when it is double
    // buffered) is probed with a non-blocking try_wait BEFORE this stage's MMAs are issued; its
    // latency overlaps the issue and the blocking wait is only entered when the data is late.
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(BLOCK_M, ACC_N);
    ptx::mbar_wait(qready_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (tile0 < tile1) {
      ptx::mbar_wait(full_bar(0), 0);
      ptx::mbar_wait(tempty_bar(0), 1u);
    }
    ptx::tc_fence_after();
    for (int tile = tile0; tile < tile1; ++tile, ++it) {
      const int acc = acc_of(it);
      const uint32_t d_tmem = tmem_base + acc * ACC_N;
      const bool last_tile = tile + 1 == tile1;
      for (int ks = 0; ks < n_kstages; ++ks) {
        const bool last_ks = ks + 1 == n_kstages;
        const int nstage = (stage + 1 == S) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == S) ? phase ^ 1u : phase;
        const bool has_next = !(last_tile && last_ks);
        const bool next_ready = has_next ? ptx::mbar_try_wait(full_bar(nstage), nphase) : true;
        const int nacc = acc_of(it + 1);
        const uint32_t nacc_par = acc_par(it + 1) ^ 1u;
        const bool probe_acc = NACC == 2 && last_ks && !last_tile;
        const bool acc_ready = probe_acc ? ptx::mbar_try_wait(tempty_bar(nacc), nacc_par) : !(last_ks && !last_tile);

        const int nk = min(STAGE_KCH, n_kch - ks * STAGE_KCH);
        const uint32_t sbase = base + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint32_t a_tmem0 = tmem_base + Q_COL0 + ks * (STAGE_KCH * KCH / 2);
        const uint64_t bdesc0 = ptx::smem_desc_sw128(sbase);
        if (ptx::elect_one()) {
#pragma unroll
          for (int c = 0; c < STAGE_KCH; ++c) {
            if (c < nk) {
#pragma unroll
              for (int j = 0; j < KCH / 16; ++j) {
                // 16 k per instruction: 8 TMEM columns of A, 32 bytes along the swizzled row of B
                ptx::mma_bf16_ts(d_tmem, a_tmem0 + c * (KCH / 2) + j * 8,
                                 bdesc0 + static_cast<uint64_t>(c * (BOX_BYTES >> 4) + 2 * j), idesc,
                                 (ks | c | j) != 0 ? 1u : 0u);
              }
            }
          }
          ptx::mma_commit(empty_bar(stage));              // frees the stage once these MMAs have read it
          if (last_ks) ptx::mma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (!next_ready) ptx::mbar_wait(full_bar(nstage), nphase);
        if (!acc_ready) ptx::mbar_wait(tempty_bar(nacc), nacc_par);
        ptx::tc_fence_after();
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== epilogue warps (thread <-> query) =====================
    const int row = warp * 32 + lane;
    const int qrow = qtile * BLOCK_M + row;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

    // query tile -> TMEM (A operand, K-major: column c of lane m holds k = 2c, 2c+1)
    {
      const uint4* qsrc = reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(qrow) * p.d_pad);
      for (int c = 0; c < p.d_pad / 16; ++c) {
        const uint4 a = qsrc[2 * c], b = qsrc[2 * c + 1];
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        ptx::tmem_st_x8(lane_addr + Q_COL0 + c * 8, v);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(qready_bar);
    }

    const int kcap = topk_kcap(p.k);
    uint32_t* set = lists + warp * 64 * kcap + lane;   // this query's keys: key i at set[i * 32], id i at set[(kcap + i) * 32]
    for (int i = 0; i < kcap; ++i) {
      set[i * TOPK_STRIDE] = i < p.k ? f32_to_ordered(-CUDART_INF_F) : 0xffffffffu;   // [k, kcap): never the worst
      set[(kcap + i) * TOPK_STRIDE] = 0xffffffffu;
    }
    int worst = 0;
    const bool live = qrow < p.nq;
    const int ign = (p.ignore_local && live) ? p.ignore_local[qrow] : -1;
    const float bkey = (p.after_key && live) ? p.after_key[qrow] : 0.f;
    const int brow = (p.after_key && live) ? p.after_row[qrow] : -1;
    float thr = -CUDART_INF_F;
    PoolState pool;   // pooling across splits is used by the CTA-pair kernel only
    pool.init(0);

    int it = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++it) {
      const int acc = acc_of(it);
      ptx::mbar_wait(tfull_bar(acc), acc_par(it));
      __syncwarp();   // tcgen05.ld is warp-collective: reconverge after the divergent admission path
      ptx::tc_fence_after();
      uint32_t v[NGRP][32];
#pragma unroll
      for (int g = 0; g < NGRP; ++g) ptx::tmem_ld_x32(lane_addr + acc * ACC_N + g * 32, v[g]);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));   // accumulator is in registers: MMA may reuse it
      fold_tile<kL2, NGRP, false>(v, p.xnorm2, tile * ACC_N, p.ntotal, ign, live, set, p.k, kcap, it == 0, thr, worst, pool,
                                      p.after_key != nullptr, bkey, brow);
    }

    if (live) {
      const size_t o = (static_cast<size_t>(split) * p.nq + qrow) * p.k;
      for (int i = 0; i < p.k; ++i) {   // unsorted: K2 merges by arg-max rounds
        p.part_key[o + i] = ordered_to_f32(set[i * TOPK_STRIDE]);
        p.part_ids[o + i] = static_cast<int>(set[(kcap + i) * TOPK_STRIDE]);
      }
    }
  }

  // teardown: every MMA has completed (the epilogue consumed the last accumulator)
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}
}  // namespace tc
