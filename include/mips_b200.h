/*
 * mips_b200.h — C ABI of the B200-native exact MIPS index (libmips_b200.so).
 *
 * This is the drop-in boundary for ONE path of florianbaud/retrieval-augmented-mds:
 * the flat exact index behind `Mips.search` (reference sotasum/mips.py:382-400), i.e. what
 * the reference reaches through HF-datasets' FaissIndex -> faiss.IndexFlat{IP,L2}
 * (call sites sotasum/mips.py:333-340 add, :383-386 search; retriever_lightning.py:317-321,
 * :400-404; pretrain.py:475-479, :519-523), plus the build-side row statistics of
 * `Mips.build_index` (mips.py:290-345) and the doc-score arithmetic of
 * retriever_generator.py:158-193.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a raw host or device address.
 *   - return 0 on success, negative MIPS_E_* on error; message via mips_last_error()
 *     (thread local). Nothing throws across the ABI.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). All device
 *     work is enqueued on it; calls do not synchronise unless documented ("sync").
 *   - ids are int64 like faiss; missing results are id -1 with score -inf (IP) / +inf (L2).
 *   - one index per process/GPU; add and search on one index must not overlap.
 */
#ifndef MIPS_B200_H_
#define MIPS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mips_index_s* mips_handle;

/* faiss.METRIC_INNER_PRODUCT / faiss.METRIC_L2 (reference mips.py:306,316,369,371). */
#define MIPS_METRIC_IP 0
#define MIPS_METRIC_L2 1

/* Storage / arithmetic type of the bank shard in HBM. */
#define MIPS_DTYPE_F32 0   /* fp32 rows (+ a bf16 shadow): EXACT fp32 search — tcgen05 filter over the shadow,
                              fp32 re-rank and certificate (MIPS_ALGO_TCX), SIMT fp32 FMA kernel as fallback */
#define MIPS_DTYPE_BF16 1  /* bf16 rows, tcgen05 tensor-core search, fp32 accumulate */

/* Search kernel selection (testing / bisection; AUTO is what the product uses). */
#define MIPS_ALGO_AUTO 0
#define MIPS_ALGO_SIMT 1   /* fp32 FMA tiled kernel, any dtype                     */
#define MIPS_ALGO_TC 2     /* tcgen05/TMEM/TMA kernel, bf16 bank, d_pad <= 768:
                              2 x 64-row double-buffered TMEM accumulators (default) */
#define MIPS_ALGO_TC128 3  /* same kernel, one 128-row accumulator (A/B comparison) */
#define MIPS_ALGO_TCX 5    /* fp32 bank, d_pad <= 1024: EXACT search at tensor-core speed — TC2 over
                              a bf16 shadow of the rows keeps kc > k candidates, their keys are recomputed
                              from the fp32 rows, a rigorous error bound certifies the top-k, and queries
                              that fail it (ties at the boundary) are recomputed by SIMT. AUTO on fp32. */
#define MIPS_ALGO_TC2 4    /* CTA-pair kernel (tcgen05 cta_group::2, M=256 x N=128), bf16 bank,
                              d_pad <= 1024: 2 x 128-row accumulators, bank tile shared by the pair */

/* Output transform applied by the merge kernel to the ranking key. */
#define MIPS_OUT_IP 0      /* D = <q,x>                    descending (IndexFlatIP)          */
#define MIPS_OUT_L2 1      /* D = |q|^2+|x|^2-2<q,x>       ascending  (IndexFlatL2)          */
#define MIPS_OUT_AUGL2 2   /* D = |q|^2+phi-2<q,x>         ascending  (IndexFlatL2 on the
                              augment_xb/augment_xq vectors of mips.py:55-70, no extra column) */

#define MIPS_E_INVALID -1  /* bad argument / shape / dtype    */
#define MIPS_E_CUDA -2     /* CUDA runtime or driver error    */
#define MIPS_E_NOMEM -3    /* allocation failed               */
#define MIPS_E_UNSUPPORTED -4
#define MIPS_E_NCCL -5     /* NCCL missing or a collective failed */

#define MIPS_MAX_K 64      /* per-pass top-k capacity of the search kernels */
#define MIPS_MAX_K_MULTIPASS 2048   /* largest k of the multi-pass entry points (mips_search_host; passes of MIPS_MAX_K
                                       through mips_search_local_after) */

/* ---- lifecycle ------------------------------------------------------------------------- */

/* Replaces faiss.index_factory(d, "Flat", metric) / IndexFlatIP(d) / IndexFlatL2(d)
 * (reference mips.py:333-340 via datasets.add_faiss_index; mips.py:665,670).
 * capacity_rows > 0 pre-sizes the HBM shard (no regrowth while ntotal <= capacity). */
int mips_create(mips_handle* out, int d, int metric, int dtype, int device, int64_t capacity_rows);
int mips_destroy(mips_handle h);
/* faiss Index.reset(): drop all rows, keep the allocation. mips_reset waits for the device first (searches
 * of this index still in flight on any stream finish before its statistics are cleared); mips_reset_async
 * is ordered on `stream` only (the double-buffered refresh resets the back shard on its side stream). */
int mips_reset(mips_handle h);
int mips_reset_async(mips_handle h, void* stream);
int64_t mips_ntotal(mips_handle h);
/* rows the HBM shard can hold without reallocating (double-buffered refresh reuses allocations) */
int64_t mips_capacity(mips_handle h);
int mips_dim(mips_handle h);
int mips_metric(mips_handle h);
int mips_dtype(mips_handle h);

/* ---- build side (K0) -------------------------------------------------------------------- */

/* Replaces faiss Index.add(x) as driven by datasets' add_vectors in 1000-row batches
 * (reference mips.py:333-340) fused with the per-row passes of Mips.build_index:
 * _map_norm (mips.py:347-349), _map_normalize (:358-361, when normalize != 0) and the
 * row |x|^2 needed by get_phi/augment_xb (:55-65). x: fp32 [n, d] row-major, host
 * (x_on_device = 0; staged through pinned memory, sync) or device (x_on_device = 1, async). */
int mips_add(mips_handle h, const float* x, int64_t n, int x_on_device, int normalize, void* stream);

/* max_i |x_i|^2 over the rows as given to mips_add (before normalisation) = get_phi(xb)
 * (mips.py:55-56) and max_norm^2 (mips.py:298-304). Sync on `stream`. */
int mips_max_norm2(mips_handle h, float* out, void* stream);
/* phi used by MIPS_OUT_AUGL2; set after the cross-rank MAX. */
int mips_set_phi(mips_handle h, float phi);
float mips_get_phi(mips_handle h);

/* faiss.normalize_L2(x) (reference mips.py:521-525): in-place row normalisation of fp32 [n, d]
 * (host: staged to the device and back, sync; device: async). Rows with zero norm untouched. */
int mips_normalize_l2(float* x, int64_t n, int d, int x_on_device, int device, void* stream);

/* Copy rows [row0, row0+n) back as fp32 [n, d] (host or device). Used by save()/np_search. */
int mips_reconstruct(mips_handle h, int64_t row0, int64_t n, float* out, int out_on_device, void* stream);

/* ---- search side (K1 + K2) -------------------------------------------------------------- */

/* Local (one shard) exact top-k. Replaces faiss Index.search(xq, k) for this shard
 * (reference mips.py:383-386) with the ignore filter of mips.py:388-398 fused in
 * ("exclude id == ignore_ids[j] for query j", ids are GLOBAL = id_offset + local row).
 *   q            device fp32 [nq, d]
 *   q_normalize  != 0: L2-normalise each query first (_prepare_query, mips.py:368-370)
 *   ignore_ids   device int64 [nq] or NULL
 *   out_key      device fp32 [nq, k]  ranking key, descending: <q,x> (IP) or <q,x>-|x|^2/2 (L2)
 *   out_ids      device int64 [nq, k] global ids, -1 padded
 *   out_xnorm2   device fp32 [nq, k] or NULL: |x|^2 of each hit (for cosine / L2 output)
 *   out_qnorm2   device fp32 [nq] or NULL: |q|^2 (after optional normalisation)
 * 1 <= k <= MIPS_MAX_K. Async on `stream`. */
int mips_search_local(mips_handle h, const float* q, int nq, int k, int q_normalize,
                      const int64_t* ignore_ids, int64_t id_offset, int algo,
                      float* out_key, int64_t* out_ids, float* out_xnorm2, float* out_qnorm2,
                      void* stream);

/* One pass of a MULTI-PASS search (k > MIPS_MAX_K; faiss accepts any k, reference mips.py:383-386): like
 * mips_search_local / mips_search_local_packed, but only rows strictly AFTER (after_key[j], after_id[j]) in the
 * total order (ranking key descending, global id ascending) are eligible for query j — pass the last result of
 * the previous pass (its out_key and out_ids; a query that ran out of rows: key -inf, id INT64_MAX). after_key /
 * after_id device [nq], both NULL = unbounded. Give out_key / out_ids (+ optional out_xnorm2) or out_packed.
 * fp32 banks run bounded passes on the exact fp32 FMA kernel. 1 <= k <= MIPS_MAX_K per pass. Async on `stream`. */
int mips_search_local_after(mips_handle h, const float* q, int nq, int k, int q_normalize, const int64_t* ignore_ids,
                            int64_t id_offset, int algo, const float* after_key, const int64_t* after_id,
                            float* out_key, int64_t* out_ids, float* out_xnorm2, float* out_qnorm2, void* out_packed,
                            void* stream);

/* k-way merge of n_parts candidate lists per query (after the cross-GPU all-gather, or of a
 * single list) + output transform + the doc-score arithmetic of retriever_generator.py:158-193.
 *   cand_key/cand_ids/cand_xnorm2  device [n_parts, nq, k_in]   (cand_xnorm2 may be NULL if
 *                                  neither L2 output nor cosine is requested)
 *   q_norm2      device [nq] or NULL (required for L2/AUGL2/cosine)
 *   D, I         device [nq, k_out] final scores (per out_mode) and ids
 *   cosine       device [nq, k_out] or NULL:  <q,x>/(|q||x|)   (retriever_generator.py:159-172)
 *   doc_prob     device [nq, k_out] or NULL:  softmax_j(beta*cosine_j + beta_bias)
 *                (the per-document factor of decoder_own.py:110-114,134)
 *   memory_bias  device [nq, k_out*mem_len] or NULL: cosine broadcast over each doc's tokens
 *                (retriever_generator.py:188-192)
 * Async on `stream`. */
int mips_merge(const float* cand_key, const int64_t* cand_ids, const float* cand_xnorm2,
               int n_parts, int nq, int k_in, int k_out, int metric, int out_mode, float phi,
               const float* q_norm2, const int64_t* ignore_ids,
               float* D, int64_t* I, float* cosine, float* doc_prob, float beta, float beta_bias,
               float* memory_bias, int mem_len, void* stream);

/* Packed variants for the multi-GPU path: candidates travel as 16-byte records
 * {float key; float xnorm2; int64 id} so that ONE all-gather moves a rank's [nq, k] list and no
 * pack/unpack kernels are needed around it (ncclAllGather of per-rank top-k, then K2).
 *   out_packed   device [nq, k] records            cand_packed  device [n_parts, nq, k_in] records */
int mips_search_local_packed(mips_handle h, const float* q, int nq, int k, int q_normalize,
                             const int64_t* ignore_ids, int64_t id_offset, int algo, void* out_packed,
                             float* out_qnorm2, void* stream);
int mips_merge_packed(const void* cand_packed, int n_parts, int nq, int k_in, int k_out, int metric,
                      int out_mode, float phi, const float* q_norm2, const int64_t* ignore_ids, float* D,
                      int64_t* I, float* cosine, float* doc_prob, float beta, float beta_bias,
                      float* memory_bias, int mem_len, void* stream);

/* End-to-end host call: what `Mips.search` does today with numpy in / numpy out
 * (reference mips.py:382-400). H2D of queries, K1, K2, D2H of (D, I); sync.
 * xq host fp32 [nq, d]; ignore_ids host int64 [nq] or NULL; D host fp32 [nq,k]; I host int64 [nq,k].
 * 1 <= k <= MIPS_MAX_K_MULTIPASS: k > MIPS_MAX_K runs ceil(k / MIPS_MAX_K) bounded passes (exact). */
int mips_search_host(mips_handle h, const float* xq, int nq, int k, int q_normalize,
                     const int64_t* ignore_ids, int out_mode, float* D, int64_t* I, void* stream);

/* ---- introspection ------------------------------------------------------------------------ */
const char* mips_last_error(void);
/* Number of kernels this library has launched so far in this process (bench gpu_launches). */
int64_t mips_launch_count(void);
/* Name of the search path the last mips_search_local call used: "tc2" (CTA-pair tcgen05), "tc" / "tc128"
 * (1-CTA tcgen05), "tcx" (exact fp32: tcgen05 filter + re-rank + certificate), "simt", or "none". */
const char* mips_last_algo(mips_handle h);
/* MIPS_ALGO_TCX statistics: queries (since the last reset) whose exactness certificate failed and
 * that were recomputed by the SIMT kernel. Sync. */
int64_t mips_fallback_queries(mips_handle h, int reset);
/* K1 timing: with profiling on, every search records a CUDA-event pair around its K1 launch on
 * the caller's stream (no sync). mips_k1_ms_total() synchronises on those events and returns the
 * SUM of the K1 durations (ms) recorded since mips_set_profiling(h, 1) (at most 256 launches are
 * kept); mips_prof_count() says how many launches that sum covers. */
int mips_set_profiling(mips_handle h, int on);
float mips_k1_ms_total(mips_handle h);
int mips_prof_count(mips_handle h);

/* ---- retrieval metrics (next row N5) ------------------------------------------------------ */

/* Replaces retriever_metrics (reference sotasum/pretrain.py:69-85; copy retriever_lightning.py:71-87)
 * together with the host loop that builds its hit matrix (sotasum/mips.py:456-463):
 * pred[b, j] = (row_aid[ids[b, j]] == query_aid[b]) (ids < 0 never hit), then
 *   out3[0] = mean_b(sum_j pred / counts[b])                         recall
 *   out3[1] = mean_b(1 / argmax_j pred, inf -> 0)                    reciprocal_rank (the reference's
 *             own definition: a first hit at index 0 scores 0, like a row without hits)
 *   out3[2] = mean_b(sum_j (cumsum(pred)_j / (j+1)) * pred_j / counts[b])   average_precision
 * All pointers are device memory; per_query [nq, 3] is scratch that also returns the per-query
 * terms; pred_out [nq, k] is optional. k <= MIPS_MAX_K_MULTIPASS. Async on `stream`. */
int mips_retriever_metrics(const int64_t* ids, int nq, int k, const int64_t* row_aid, int64_t n_rows,
                           const int64_t* query_aid, const float* counts, float* per_query, float* out3,
                           float* pred_out, void* stream);

/* ---- result gather for the consumer (next row N2) ------------------------------------------ */

/* Stored rows of the shard by GLOBAL id as fp32 [n, d] (device): replaces the host gather
 * `self.embeddings[i]` + re-encoding of reference sotasum/mips.py:428,465-470 when the memory encoder is
 * frozen, so that the cosine doc score of retriever_generator.py:158-172 can be recomputed with gradient
 * w.r.t. the query. ids outside [id_offset, id_offset + ntotal) (other shards, -1 padding) give zero
 * rows: a row-sharded bank sums the ranks' outputs. Async on `stream`. */
int mips_gather_rows(mips_handle h, const int64_t* ids, int64_t n, int64_t id_offset, float* out, void* stream);

/* Pre-tokenised memory store [n_rows, L] int32 (+ token counts [n_rows]) in HBM -> the four tensors the
 * reference builds per step by re-tokenising the retrieved texts on the host (sotasum/mips.py:473-501):
 * input_ids, attention_mask (t < len), memory_attention_mask (attention_mask with bos/eos positions
 * cleared; optional) and global_attention_mask (1 on the first token; optional), all int64 [n, L].
 * ids < 0 or >= n_rows give pad rows with zero masks. All pointers device memory. Async on `stream`. */
int mips_gather_tokens(const int32_t* store_ids, const int32_t* store_len, int64_t n_rows, int L, const int64_t* ids,
                       int64_t n, int32_t pad_id, int32_t bos_id, int32_t eos_id, int64_t* input_ids,
                       int64_t* attention_mask, int64_t* memory_attention_mask, int64_t* global_attention_mask,
                       void* stream);

/* ---- peer-memory exchange of the per-rank lists (multi-GPU, one box) ------------------------ */

/* The cross-GPU step of the sharded search (absent in the reference, which never shards) without a
 * collective launch: every rank owns an exchange buffer that its peers map through CUDA IPC; the local
 * merge kernel of rank r stores its [nq, k] 16-byte records into slot r of EVERY rank's buffer over
 * NVLink and raises rank r's flag there; the final merge kernel of each rank waits for the G flags of
 * this search (`seq`) and merges its own buffer. Callers alternate two slot sets between consecutive
 * searches (a rank may be one search ahead of a peer).
 *   mips_xchg_alloc : cudaMalloc + zero + IPC handle (64 bytes) of a buffer on `device`
 *   mips_xchg_open  : map a peer's buffer from its handle; mips_xchg_close / mips_xchg_free undo them
 *   mips_search_local_xchg : mips_search_local_packed whose output goes to peer_bufs[g] (device array of
 *       n_peers pointers to this rank's [nq, k] region on each rank, itself included) and whose
 *       completion is signalled on peer_flags[g]; bf16 / SIMT single-chunk searches only
 *   mips_merge_xchg : mips_merge_packed over my_buf = [n_ranks, nq, k_in] records once my_flags[0..n_ranks)
 *       all read `seq`. my_flags is a block of 64 uint32: [0, 32) arrival flags, word MIPS_XCHG_TIMEOUT_WORD
 *       the time-out flag. The wait is bounded in wall time (MIPS_XCHG_TIMEOUT_S, default 120 s); when it
 *       expires the affected queries return ids -1 and the time-out flag receives `seq` — the context stays
 *       usable and the caller falls back to the NCCL exchange
 *   mips_xchg_timeout_seq : read that flag (0 = no search timed out so far). Sync. */
#define MIPS_XCHG_TIMEOUT_WORD 32
int mips_xchg_alloc(int device, int64_t bytes, void** ptr, void* handle64);
int mips_xchg_open(int device, const void* handle64, void** ptr);
int mips_xchg_close(int device, void* ptr);
int mips_xchg_free(int device, void* ptr);
int mips_search_local_xchg(mips_handle h, const float* q, int nq, int k, int q_normalize, const int64_t* ignore_ids,
                           int64_t id_offset, int algo, void* const* peer_bufs, uint32_t* const* peer_flags,
                           int n_peers, uint32_t seq, float* out_qnorm2, void* stream);
int mips_merge_xchg(const void* my_buf, const uint32_t* my_flags, int n_ranks, uint32_t seq, int nq, int k_in, int k_out,
                    int metric, int out_mode, float phi, const float* q_norm2, const int64_t* ignore_ids, float* D,
                    int64_t* I, float* cosine, float* doc_prob, float beta, float beta_bias, float* memory_bias,
                    int mem_len, void* stream);
int mips_xchg_timeout_seq(int device, const uint32_t* my_flags, uint32_t* out_seq);

/* ---- cross-GPU step through NCCL (K3) --------------------------------------------------------- */

/* The reference never shards (rank 0 builds, every rank loads a full CPU replica: lightning_model.py:168-180);
 * here the bank is row-sharded with the partition of encode_text2 (sotasum/mips.py:226-230) and each search
 * combines the per-rank lists with ONE collective on the caller's stream. libnccl.so.2 is
 * bound at run time (the copy the process already loaded, e.g. the one bundled with the host's tensor
 * framework; else the system one; MIPS_NCCL_LIB
 * overrides) — `nccl_comm` is an ncclComm_t passed as void*, either created by mips_nccl_comm_init or any
 * communicator of the SAME libnccl the host already owns.
 *   mips_nccl_version     : NCCL_VERSION_CODE of the bound library, -1 when NCCL is unavailable
 *   mips_nccl_unique_id   : ncclGetUniqueId into id128 (128 bytes) — rank 0, then broadcast by the host
 *   mips_nccl_comm_init   : ncclCommInitRank on `device` (collective over the n_ranks processes)
 *   mips_allgather_topk   : ncclAllGather of this rank's [nq, k] 16-byte records -> [n_ranks, nq, k]
 *   mips_search_sharded   : the whole step, queries REPLICATED on every rank (same q everywhere):
 *                           query prep + K1 + local merge + all-gather + final merge with the fused doc-score
 *                           outputs of mips_merge. Every launch goes to `stream`, nothing synchronises, scratch is
 *                           owned by the handle and stable after the first call: the call can be captured in a
 *                           CUDA graph (cudaStreamBeginCapture ... EndCapture) and replayed.
 *   mips_search_sharded_dp: the data-parallel TRAINING step — each rank has its OWN nq_local queries, as every
 *                           DDP rank calls self.mips(queries=...) with its own batch (retriever_generator.py:143-153
 *                           driven per rank by lightning_model.py:188-216): all-gather the queries (and ignored
 *                           ids) -> one local search of n_ranks * nq_local queries -> all-to-all of the records
 *                           (ncclSend/ncclRecv group) -> each rank merges its own nq_local queries. nq_local and
 *                           "ignore_local is NULL" must agree across ranks. Outputs are [nq_local, ...]. */
int mips_nccl_version(void);
int mips_nccl_unique_id(void* id128);
int mips_nccl_comm_init(void** comm, int n_ranks, int rank, const void* id128, int device);
int mips_nccl_comm_destroy(void* comm);
int mips_allgather_topk(mips_handle h, void* nccl_comm, const void* local_packed, void* gathered_packed, int nq, int k,
                        void* stream);
int mips_search_sharded(mips_handle h, void* nccl_comm, int n_ranks, const float* q, int nq, int k, int q_normalize,
                        const int64_t* ignore_ids, int64_t id_offset, int algo, int out_mode, float* D, int64_t* I,
                        float* cosine, float* doc_prob, float beta, float beta_bias, float* memory_bias, int mem_len,
                        void* stream);
int mips_search_sharded_dp(mips_handle h, void* nccl_comm, int n_ranks, int rank, const float* q_local, int nq_local,
                           int k, int q_normalize, const int64_t* ignore_local, int64_t id_offset, int algo,
                           int out_mode, float* D, int64_t* I, float* cosine, float* doc_prob, float beta,
                           float beta_bias, float* memory_bias, int mem_len, void* stream);

/* ---- generation / copy mixture (next row N3) ------------------------------------------------- */

/* Replaces reference sotasum/retriever_generator.py:391-404:
 *   out[r, :] = log(gen_gate[r] * softmax(logits[r, :]) + scatter_add(copy_probs[r, :] at copy_seq[b, :]) + eps)
 * with r = b * rows_per_batch + t, logits / out fp32 [n_rows, V], gen_gate [n_rows], copy_probs [n_rows, S]
 * (already gated: copy_gate * attention, decoder_own.py:538), copy_seq int64 [n_rows / rows_per_batch, S]
 * (tokens outside [0, V) are skipped), eps = 1e-7 in the reference. One pass over the logits, the
 * vocabulary row lives in shared memory: V <= 56000. All pointers device memory; async on `stream`.
 *   mips_copy_mixture      forward (generation / evaluation)
 *   mips_copy_mixture_fwd  forward that also saves the row statistics (stats [n_rows, 2]: max and
 *                          sum of exp(logits - max)) for the backward; stats may be NULL
 *   mips_copy_mixture_bwd  training: for the upstream gradient dout [n_rows, V] writes dlogits [n_rows, V],
 *                          dgate [n_rows] and dcopy [n_rows, S] (the gradient of copy_probs) in one pass over
 *                          (logits, out, dout) */
int mips_copy_mixture(const float* logits, const float* gen_gate, const float* copy_probs, const int64_t* copy_seq,
                      int64_t n_rows, int rows_per_batch, int V, int S, float eps, float* out, void* stream);
int mips_copy_mixture_fwd(const float* logits, const float* gen_gate, const float* copy_probs, const int64_t* copy_seq,
                          int64_t n_rows, int rows_per_batch, int V, int S, float eps, float* out, float* stats,
                          void* stream);
int mips_copy_mixture_bwd(const float* logits, const float* out, const float* dout, const float* gen_gate,
                          const float* stats, const int64_t* copy_seq, int64_t n_rows, int rows_per_batch, int V, int S,
                          float* dlogits, float* dgate, float* dcopy, void* stream);

/* ---- score-biased copy attention (next row N3) ------------------------------------------------ */

/* The elementwise / row-wise part of the copy decoder's ONE-head cross attention, reference
 * sotasum/decoder_own.py:110-134 (the GEMMs at :108 and :164 stay library GEMMs of the host):
 *   probs[r, s] = softmax_s( scores[r, s] + (beta * doc_scores[b, s / mem_len] + beta_bias) + mask[b, s] )
 * r = b * T + t; scores / probs fp32 [B*T, S]; doc_scores [B, n_docs] = `mips_scores` — its broadcast
 * `memory_bias` (retriever_generator.py:188-192) is never materialised (mem_len = 1, n_docs = S gives a general
 * per-token attention_bias); mask additive [B, S] (0 / finfo.min) or NULL; beta_dev = device {beta, beta_bias}
 * (trainable parameters; read on the device, no host sync) or NULL to use the two scalars. S <= 16384.
 *   _fwd: one pass (read scores, write probs; probs may alias scores)
 *   _bwd: dscores = probs * (dprobs - sum_s probs * dprobs) and, fused, the reduction that carries the gradient
 *         to the retriever: doc_grad[b, j] += sum_t sum_{s in doc j} dscores (zeroed by the caller; NULL to skip).
 *         d doc_scores = beta * doc_grad, d beta = sum(doc_grad * doc_scores), d beta_bias = sum(doc_grad). */
int mips_copy_attention_softmax_fwd(const float* scores, const float* doc_scores, int n_docs, int mem_len, float beta,
                                    float beta_bias, const float* beta_dev, const float* mask, int B, int T, int S,
                                    float* probs, void* stream);
int mips_copy_attention_softmax_bwd(const float* probs, const float* dprobs, int n_docs, int mem_len, int B, int T, int S,
                                    float* dscores, float* doc_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIPS_B200_H_ */
