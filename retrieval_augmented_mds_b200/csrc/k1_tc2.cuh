// k1_tc2.cuh — K1 (CTA-pair tensor-core variant): the query x bank contraction of
// faiss.IndexFlat*.search as reached from Mips.search (reference sotasum/mips.py:383-386) and of
// inner_product (mips.py:552-560), as ONE tcgen05.mma.cta_group::2 stream per SM pair with the
// fused per-query top-k epilogue of k1_tc.cuh.
//
// Why a pair (measured on the 1-CTA kernel, profiles/r01_k1_tc_metrics.json): with 8 query tiles
// re-reading every bank tile from L2 the kernel moved 123 GB L2->SM per launch (9.4 TB/s, the
// crossbar ceiling), and its 64-column accumulators cap the tensor pipe at 73.5 % (an MMA costs
// N/2 + ~11 cycles). A CTA pair shares every bank tile: each CTA fetches HALF of it (L2->SM
// traffic halves) and the instruction becomes M=256 x N=128 (86.5 % ceiling).
//
// Mapping (cluster = 2 CTAs on one TPC, persistent over a slice of the bank):
//   * 256 queries per pair, 128 per CTA = the 128 TMEM lanes of that CTA. The query tile is
//     STATIONARY for the whole kernel. TMEM has 512 columns: [0,256) hold two 128-column fp32
//     accumulators (MMA of tile t+1 overlaps the epilogue of tile t), [256,512) hold the first
//     512 dims of the query tile as packed bf16 pairs (A operand from TMEM); dims >= 512 stay in
//     shared memory (A operand from a shared-memory descriptor, loaded once by TMA).
//   * a bank tile is 128 rows; CTA r streams rows [128 t + 64 r, +64) by TMA (64 x 64 boxes,
//     128-byte swizzle, SKCH boxes per stage, mbarrier ring). Both CTAs' TMA loads credit the
//     LEADER's full barrier; tcgen05.commit multicasts "stage free" / "accumulator full" to both.
//   * the leader's MMA warp issues 4 x (d_pad/64) instructions per tile (M=256, N=128, K=16).
//   * epilogue (both CTAs, 4 warps each): thread <-> query as in k1_tc.cuh; the accumulator is
//     handed back to the leader with a remote mbarrier arrive.
//
// Grid = 2 * n_qpairs * n_splits CTAs (<= #SMs). Roofline: tensor bound,
// 2*256*128*d_pad flops per pair tile; HBM bytes = one pass over the bank per <= 256*n_qpairs
// queries; L2->SM bytes = n_qpairs passes (was 2*n_qpairs).
#pragma once
#include "common.cuh"
#include "k1_topk.cuh"
#include "ptx.cuh"

namespace tc2 {
constexpr int BLOCK_M = 128;        // queries per CTA (TMEM lanes)
constexpr int PAIR_M = 256;         // queries per CTA pair (MMA M)
constexpr int TILE_N = 128;         // bank rows per pair tile (MMA N)
constexpr int HALF_N = 64;          // bank rows each CTA loads per tile
constexpr int KCH = 64;             // bf16 per 128-byte swizzle row
constexpr int TMEM_COLS = 512;
constexpr int Q_COL0 = 2 * TILE_N;  // query tile starts after the two accumulators
constexpr int TMEM_KCH = (TMEM_COLS - Q_COL0) * 2 / KCH;   // 8 k-chunks (512 dims) of A in TMEM
constexpr int MAX_KCH = 16;         // d_pad <= 1024
constexpr int THREADS = 192;        // 4 epilogue warps + TMA warp + MMA warp
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;  // 227 KiB opt-in maximum
constexpr int BBOX_BYTES = HALF_N * KCH * 2;    // 8 KiB: one bank box (64 rows x 64 k)
constexpr int ABOX_BYTES = BLOCK_M * KCH * 2;   // 16 KiB: one query box (128 rows x 64 k)
constexpr int N_BARS = 2 * MAX_STAGES + 8;

// shared memory: [align pad 1024][A tail boxes][stages][top-k sets][barriers]
__host__ __device__ inline int a_smem_kch(int d_pad) {
  const int n = d_pad / KCH - TMEM_KCH;
  return n > 0 ? n : 0;
}
__host__ __device__ inline int list_bytes(int k) { return BLOCK_M * topk_kcap(k) * 8; }
inline int pick_stages(int d_pad, int k, int skch) {
  const int avail = SMEM_LIMIT - 1024 - a_smem_kch(d_pad) * ABOX_BYTES - list_bytes(k) - N_BARS * 8;
  int s = avail / (skch * BBOX_BYTES);
  return s > MAX_STAGES ? MAX_STAGES : s;
}
inline size_t smem_bytes(int d_pad, int k, int stages, int skch) {
  return 1024 + static_cast<size_t>(a_smem_kch(d_pad)) * ABOX_BYTES +
         static_cast<size_t>(stages) * skch * BBOX_BYTES + list_bytes(k) + N_BARS * 8;
}

struct Params {
  const __nv_bfloat16* q;   // [n_qpairs*256, d_pad] prepared queries (zero padded)
  const float* xnorm2;      // [capacity] (L2 only)
  const int* ignore_local;  // [nq] or null
  const float* after_key;   // [nq] or null: multi-pass search, only rows strictly after (after_key, after_row) ...
  const int* after_row;     // ... in the order (key descending, row ascending) are eligible
  float* part_key;          // [n_splits, nq, k]
  int* part_ids;
  int64_t ntotal;
  int nq, d_pad, k, n_tiles, n_qpairs, n_splits, stages;
  unsigned long long cache_hint;
  int* pace;        // [n_splits, n_qpairs] tiles issued so far by each pair (zeroed before the launch), or null
  int pace_window;  // a pair never runs more than this many tiles ahead of the slowest pair of its split
  uint32_t* pool;   // [n_qpairs * 256, n_splits] ordered image of every split's m-th best admitted key per query
                    // (zeroed before the launch = not published); kPool kernels only: pooled admission threshold, k1_topk.cuh
  int pool_m;       // m = ceil(k / n_splits) in [1, 4]
};

template <bool kL2, int SKCH, bool kPool>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) search_tc2_kernel(
    const __grid_constant__ CUtensorMap tmap_bank, const __grid_constant__ CUtensorMap tmap_q,
    const Params p) {
  constexpr int STAGE_BYTES = SKCH * BBOX_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;   // 128B swizzle atoms need 1024B alignment
  uint8_t* gen = smem_raw + (base - raw_addr);

  const int S = p.stages;
  const int n_kch = p.d_pad / KCH;
  const int n_akch = a_smem_kch(p.d_pad);               // query k-chunks kept in shared memory
  const uint32_t a_off = 0;
  const uint32_t st_off = static_cast<uint32_t>(n_akch) * ABOX_BYTES;
  const uint32_t lists_off = st_off + static_cast<uint32_t>(S) * STAGE_BYTES;
  uint32_t* lists = reinterpret_cast<uint32_t*>(gen + lists_off);   // per warp: keys [kcap][32], ids [kcap][32]
  const uint32_t bars_off = lists_off + list_bytes(p.k);
  const uint32_t bars = base + bars_off;
  auto full_bar = [&](int i) { return bars + 8u * i; };                       // leader's is used
  auto empty_bar = [&](int i) { return bars + 8u * (MAX_STAGES + i); };       // one per CTA
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * MAX_STAGES + a); };   // one per CTA
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * MAX_STAGES + 2 + a); };  // leader's
  const uint32_t qready_bar = bars + 8u * (2 * MAX_STAGES + 4);               // leader's
  const uint32_t qsmem_bar = bars + 8u * (2 * MAX_STAGES + 5);                // leader's
  const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 6);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + bars_off + 8 * (2 * MAX_STAGES + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int qpair = pair % p.n_qpairs, split = pair / p.n_qpairs;
  const int tile0 = static_cast<int>(static_cast<int64_t>(split) * p.n_tiles / p.n_splits);
  const int tile1 = static_cast<int>(static_cast<int64_t>(split + 1) * p.n_tiles / p.n_splits);
  const int n_kstages = (n_kch + SKCH - 1) / SKCH;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_bank);
    ptx::prefetch_tensormap(&tmap_q);
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(full_bar(i), 1);    // leader producer's arrive.expect_tx (bytes of both CTAs)
      ptx::mbar_init(empty_bar(i), 1);   // tcgen05.commit (multicast)
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);    // tcgen05.commit (multicast)
      ptx::mbar_init(tempty_bar(a), 8);   // one lane per epilogue warp of both CTAs
    }
    ptx::mbar_init(qready_bar, 8);        // query tile resident in TMEM: 4 warps x 2 CTAs
    ptx::mbar_init(qsmem_bar, 1);         // query tail resident in shared memory (both CTAs)
    ptx::fence_mbar_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc_pair(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // barriers of BOTH CTAs are initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  auto acc_par = [](int it) { return static_cast<uint32_t>((it >> 1) & 1); };

  if (warp == 4) {
    // ===================== TMA producer (both CTAs) =====================
    const uint32_t full0_leader = ptx::mapa(full_bar(0), 0);
    const uint32_t qsmem_leader = ptx::mapa(qsmem_bar, 0);
    if (n_akch > 0) {
      if (ptx::elect_one()) {
        if (rank == 0)
          ptx::mbar_arrive_expect_tx(qsmem_bar, 2u * static_cast<uint32_t>(n_akch) * ABOX_BYTES);
        for (int c = 0; c < n_akch; ++c)
          ptx::tma_load_2d_pair_hint(base + a_off + c * ABOX_BYTES, &tmap_q, qsmem_leader,
                                     (TMEM_KCH + c) * KCH,
                                     qpair * PAIR_M + static_cast<int>(rank) * BLOCK_M,
                                     ptx::kEvictNormal);
      }
      __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    // Lock-step pacing. The n_qpairs pairs of one split stream the SAME bank tiles; nothing else
    // keeps them together, and once they drift apart by more than L2 holds (126 MB over 18 splits
    // = ~36 tiles) every pair fetches its tiles from HBM again (ncu: 40 GB read per launch for a
    // 15.4 GB bank, L2 hit rate 41 %). The leader's producer publishes its tile count and does not
    // run more than `pace_window` tiles ahead of the slowest pair of its split. The wait is
    // bounded: pacing is a locality hint, never a correctness dependency between CTAs.
    int* pace_mine = p.pace ? p.pace + split * p.n_qpairs + qpair : nullptr;
    const int* pace_peer = (p.pace && lane < p.n_qpairs) ? p.pace + split * p.n_qpairs + lane : nullptr;
    const bool pacing = p.pace != nullptr && rank == 0 && p.n_qpairs > 1;
    int peer_seen = 0;   // peers' tile counts, loaded while the previous tile's stages were being issued
    for (int tile = tile0; tile < tile1; ++tile) {
      const int row0 = tile * TILE_N + static_cast<int>(rank) * HALF_N;
      for (int ks = 0; ks < n_kstages; ++ks) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        if (pacing && ks == 0) {
          // the counters were requested a tile ago (their latency hid behind the wait above); only a
          // pair that is actually too far ahead pays for fresh loads
          const int done = tile - tile0;
          if (lane == 0) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(pace_mine), "r"(done) : "memory");
          int slowest = __reduce_min_sync(0xffffffffu, pace_peer ? peer_seen : 0x7fffffff);
          for (int spin = 0; done - slowest > p.pace_window && spin < 256; ++spin) {
            __nanosleep(200);
            int other = 0x7fffffff;
            if (pace_peer) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(other) : "l"(pace_peer) : "memory");
            slowest = __reduce_min_sync(0xffffffffu, other);
          }
        }
        const int nk = min(SKCH, n_kch - ks * SKCH);
        const uint32_t dst = base + st_off + static_cast<uint32_t>(stage) * STAGE_BYTES;
        if (ptx::elect_one()) {
          if (rank == 0)
            ptx::mbar_arrive_expect_tx(full_bar(stage), 2u * static_cast<uint32_t>(nk) * BBOX_BYTES);
#pragma unroll
          for (int c = 0; c < SKCH; ++c)
            if (c < nk)
              ptx::tma_load_2d_pair_hint(dst + c * BBOX_BYTES, &tmap_bank, full0_leader + 8u * stage,
                                         (ks * SKCH + c) * KCH, row0, p.cache_hint);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (pacing && pace_peer)
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(peer_seen) : "l"(pace_peer) : "memory");
    }
    if (pacing && lane == 0)
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(pace_mine), "r"(0x7fffffff) : "memory");
    // tail: every multicast "stage free" signal addressed to this CTA has landed before it exits
    for (int i = 0; i < S; ++i) {
      ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
      if (++stage == S) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 5) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(PAIR_M, TILE_N);
      ptx::mbar_wait(qready_bar, 0);
      if (n_akch > 0) ptx::mbar_wait(qsmem_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (tile0 < tile1) {
        ptx::mbar_wait(full_bar(0), 0);
        ptx::mbar_wait(tempty_bar(0), 1u);
      }
      ptx::tc_fence_after();
      for (int tile = tile0; tile < tile1; ++tile, ++it) {
        const int acc = it & 1;
        const uint32_t d_tmem = tmem_base + acc * TILE_N;
        const bool last_tile = tile + 1 == tile1;
        for (int ks = 0; ks < n_kstages; ++ks) {
          const bool last_ks = ks + 1 == n_kstages;
          const int nstage = (stage + 1 == S) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == S) ? phase ^ 1u : phase;
          const bool has_next = !(last_tile && last_ks);
          const bool next_ready = has_next ? ptx::mbar_test_wait(full_bar(nstage), nphase) : true;
          const int nacc = (it + 1) & 1;
          const uint32_t nacc_par = acc_par(it + 1) ^ 1u;
          const bool probe_acc = last_ks && !last_tile;
          const bool acc_ready = probe_acc ? ptx::mbar_test_wait(tempty_bar(nacc), nacc_par) : true;

          const int nk = min(SKCH, n_kch - ks * SKCH);
          const uint32_t sbase = base + st_off + static_cast<uint32_t>(stage) * STAGE_BYTES;
          const uint64_t bdesc0 = ptx::smem_desc_sw128(sbase);
          if (ptx::elect_one()) {
#pragma unroll
            for (int c = 0; c < SKCH; ++c) {
              if (c < nk) {
                const int kc = ks * SKCH + c;
                const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(c * (BBOX_BYTES >> 4));
                if (kc < TMEM_KCH) {
                  const uint32_t a_tmem = tmem_base + Q_COL0 + kc * (KCH / 2);
#pragma unroll
                  for (int j = 0; j < KCH / 16; ++j)
                    ptx::mma_bf16_ts_pair(d_tmem, a_tmem + j * 8, bdesc + 2 * j, idesc,
                                          (kc | j) != 0 ? 1u : 0u);
                } else {
                  const uint64_t adesc =
                      ptx::smem_desc_sw128(base + a_off + (kc - TMEM_KCH) * ABOX_BYTES);
#pragma unroll
                  for (int j = 0; j < KCH / 16; ++j)
                    ptx::mma_bf16_ss_pair(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, 1u);
                }
              }
            }
            ptx::mma_commit_pair(empty_bar(stage));              // both CTAs may refill the stage
            if (last_ks) ptx::mma_commit_pair(tfull_bar(acc));   // accumulator complete -> epilogues
          }
          __syncwarp();
          if (!next_ready) ptx::mbar_wait(full_bar(nstage), nphase);
          if (!acc_ready) ptx::mbar_wait(tempty_bar(nacc), nacc_par);
          ptx::tc_fence_after();
          stage = nstage;
          phase = nphase;
        }
      }
    }
  } else {
    // ===================== epilogue warps (thread <-> query, both CTAs) =====================
    const int row = warp * 32 + lane;
    const int qrow = qpair * PAIR_M + static_cast<int>(rank) * BLOCK_M + row;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t qready_leader = ptx::mapa(qready_bar, 0);
    const uint32_t tempty0_leader = ptx::mapa(tempty_bar(0), 0);

    // first min(d_pad, 512) dims of the query tile -> TMEM (A operand, K-major: column c of lane m
    // holds k = 2c, 2c+1)
    {
      const uint4* qsrc = reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(qrow) * p.d_pad);
      const int n_c = min(n_kch, TMEM_KCH) * (KCH / 16);
      for (int c = 0; c < n_c; ++c) {
        const uint4 a = qsrc[2 * c], b = qsrc[2 * c + 1];
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        ptx::tmem_st_x8(lane_addr + Q_COL0 + c * 8, v);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(qready_leader);
    }

    const int kcap = topk_kcap(p.k);
    uint32_t* set = lists + warp * 64 * kcap + lane;   // this query's keys: key i at set[i * 32], id i at set[(kcap + i) * 32]
    for (int i = 0; i < kcap; ++i) {
      set[i * TOPK_STRIDE] = i < p.k ? f32_to_ordered(-CUDART_INF_F) : 0xffffffffu;   // [k, kcap): never the worst
      set[(kcap + i) * TOPK_STRIDE] = 0xffffffffu;
    }
    int worst = 0;
    float thr = -CUDART_INF_F;
    const bool live = qrow < p.nq;
    const int ign = (p.ignore_local && live) ? p.ignore_local[qrow] : -1;
    const float bkey = (p.after_key && live) ? p.after_key[qrow] : 0.f;
    const int brow = (p.after_key && live) ? p.after_row[qrow] : -1;
    PoolState pool;
    pool.init(kPool ? p.pool_m : 0);
    uint32_t* pub = kPool ? p.pool + static_cast<size_t>(qrow) * p.n_splits : nullptr;
    float published = -CUDART_INF_F;

    int it = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++it) {
      const int acc = it & 1;
      if (kPool && live && (it < 8 || (it & 7) == 0)) {
        // T = min over the splits of their published m-th best (0 = not yet published: no pooling yet); the
        // loads are in flight while this warp waits for the accumulator below
        uint32_t t = 0xffffffffu;
        for (int sp = 0; sp < p.n_splits; ++sp) {
          uint32_t u;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(u) : "l"(pub + sp) : "memory");
          t = min(t, u);
        }
        if (t > 0u) {
          pool.floor = fmaxf(pool.floor, ordered_to_f32(t - 1u));
          thr = fmaxf(thr, pool.floor);
        }
      }
      ptx::mbar_wait(tfull_bar(acc), acc_par(it));
      __syncwarp();   // tcgen05.ld is warp-collective: reconverge after the divergent insert path
      ptx::tc_fence_after();
      const int id0 = tile * TILE_N;
      // the whole 128-column accumulator row goes to registers at once (the CTA has 341 registers
      // per thread to spend) so that the accumulator is handed back before any top-k work
      uint32_t v[4][32];
      ptx::tmem_ld_x32(lane_addr + acc * TILE_N, v[0]);
      ptx::tmem_ld_x32(lane_addr + acc * TILE_N + 32, v[1]);
      ptx::tmem_ld_x32(lane_addr + acc * TILE_N + 64, v[2]);
      ptx::tmem_ld_x32(lane_addr + acc * TILE_N + 96, v[3]);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty0_leader + 8u * acc);   // accumulator is in registers
      fold_tile<kL2, 4, kPool>(v, p.xnorm2, id0, p.ntotal, ign, live, set, p.k, kcap, it == 0, thr, worst, pool,
                               p.after_key != nullptr, bkey, brow);
      if (kPool && live) {
        const float mth = pool.mth();
        if (mth > published) {   // published values only grow
          published = mth;
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(pub + split), "r"(f32_to_ordered(mth)) : "memory");
        }
      }
    }

    if (live) {
      const size_t o = (static_cast<size_t>(split) * p.nq + qrow) * p.k;
      for (int i = 0; i < p.k; ++i) {   // unsorted: K2 merges by arg-max rounds
        p.part_key[o + i] = ordered_to_f32(set[i * TOPK_STRIDE]);
        p.part_ids[o + i] = static_cast<int>(set[(kcap + i) * TOPK_STRIDE]);
      }
    }
  }

  // teardown: every MMA has completed (both epilogues consumed the last accumulator) and no
  // remote signal is in flight towards a CTA that has left
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}
}  // namespace tc2
