// mips_api.cu — host side of the C ABI declared in include/mips_b200.h.
// Owns the HBM bank shard, the norm array and the scratch buffers; enqueues K0 (ingest),
// K1 (search, tcgen05 or SIMT) and K2 (merge) on the caller's stream. No CPU compute path:
// every entry point that searches launches CUDA kernels or fails.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "k0_rows.cuh"
#include "k1_simt.cuh"
#include "k1_tc.cuh"
#include "k1_tc2.cuh"
#include "k2_merge.cuh"
#include "k3_rerank.cuh"
#include "k4_metrics.cuh"
#include "k5_gather.cuh"
#include "k6_mixture.cuh"
#include "k7_attention.cuh"
#include "mips_b200.h"
#include "nccl_dl.h"

// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// tuning knobs from the environment (A/B runs only), read once per process
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return set_err(_e == cudaErrorMemoryAllocation ? MIPS_E_NOMEM : MIPS_E_CUDA, "%s: %s (%s:%d)", \
                     #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                        \
  } while (0)

#define LAUNCH_CHECK(name)                                                                   \
  do {                                                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return set_err(MIPS_E_CUDA, "launch %s: %s", name, cudaGetErrorString(_e));            \
  } while (0)

constexpr int kProfSlots = 256;

// K1 timing events: on a capturing stream the record becomes an EXTERNAL event-record node of the graph (the event
// is re-recorded by every replay and can be read with cudaEventElapsedTime afterwards)
static cudaError_t prof_record(cudaEvent_t ev, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
    return cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal);
  return cudaEventRecord(ev, st);
}

struct mips_index_s {
  int d = 0, d_pad = 0, metric = 0, dtype = 0, device = 0;
  int64_t ntotal = 0, capacity = 0;
  void* bank = nullptr;
  float* norm2 = nullptr;
  unsigned int* max_norm2_bits = nullptr;  // device scalars: [0] max |x|^2 of the RAW rows (get_phi),
                                           // [1] max |x|^2 of the STORED rows, [2] max |x - bf16(x)|^2,
                                           // [3] certificate fallback counter, [4] exchange completion counter
  // fp32 index only: bf16 shadow of the rows (tensor-core filter of the exact search, k3_rerank.cuh)
  __nv_bfloat16* shadow = nullptr;
  CUtensorMap tmap_shadow64;
  bool shadow_valid = false;
  float phi = 0.f;
  int sm_count = 148;
  // TMA descriptor of the bank (re-encoded when the allocation changes)
  CUtensorMap tmap128;   // 128-row boxes (single 128-wide accumulator kernel)
  CUtensorMap tmap64;    // 64-row boxes (double-buffered 64-wide accumulator kernel)
  bool tmap_valid = false;
  // TMA descriptor of the prepared queries (128-row boxes; CTA-pair kernel), re-encoded when the
  // scratch allocation or its padded row count changes
  CUtensorMap tmap_q;
  const void* tmap_q_ptr = nullptr;
  int tmap_q_rows = 0;
  // scratch (grown on demand; stable after warm-up)
  void* q_prep = nullptr;      size_t q_prep_bytes = 0;
  float* q_norm2 = nullptr;    size_t q_norm2_bytes = 0;
  float* part_key = nullptr;   size_t part_key_bytes = 0;
  int* part_ids = nullptr;     size_t part_ids_bytes = 0;
  int* ign_local = nullptr;    size_t ign_bytes = 0;
  int* after_row = nullptr;    size_t after_row_bytes = 0;  // multi-pass bound as shard-local rows
  float* hafter_key = nullptr; size_t hafter_key_bytes = 0; // mips_search_host, k > MIPS_MAX_K: bound of the next pass
  int64_t* hafter_id = nullptr; size_t hafter_id_bytes = 0;
  int* pace = nullptr;         size_t pace_bytes = 0;
  uint32_t* pool = nullptr;    size_t pool_bytes = 0;       // pooled admission thresholds [nq_pad, n_splits]
  __nv_bfloat16* q_hi = nullptr; size_t q_hi_bytes = 0;     // bf16-rounded prepared queries
  float* q_res2 = nullptr;     size_t q_res2_bytes = 0;     // |q - bf16(q)|^2
  float* cand_key = nullptr;   size_t cand_key_bytes = 0;   // approximate keys of the kc candidates
  int64_t* cand_rows = nullptr; size_t cand_rows_bytes = 0; // their shard-local rows
  int* fb_flags = nullptr;     size_t fb_flags_bytes = 0;   // [nq] need_fallback + [ceil(nq/64)] tile flags + [1] counter
  int64_t fallback_queries = 0;                             // statistics of the last search (host, lazily synced)
  int* fb_count_dev = nullptr;
  float* stage_x = nullptr;    size_t stage_x_bytes = 0;
  // sharded step (K3): this rank's packed list, the gathered lists, |q|^2, and (DP variant) the gathered queries
  void* sh_local = nullptr;    size_t sh_local_bytes = 0;
  void* sh_gath = nullptr;     size_t sh_gath_bytes = 0;
  float* sh_qn2 = nullptr;     size_t sh_qn2_bytes = 0;
  float* dp_q = nullptr;       size_t dp_q_bytes = 0;
  int64_t* dp_ign = nullptr;   size_t dp_ign_bytes = 0;
  // host-call scratch
  float* hq = nullptr;         size_t hq_bytes = 0;
  int64_t* hign = nullptr;     size_t hign_bytes = 0;
  float* hkey = nullptr;       size_t hkey_bytes = 0;
  int64_t* hids = nullptr;     size_t hids_bytes = 0;
  float* hxn2 = nullptr;       size_t hxn2_bytes = 0;
  float* hqn2 = nullptr;       size_t hqn2_bytes = 0;
  float* hD = nullptr;         size_t hD_bytes = 0;
  int64_t* hI = nullptr;       size_t hI_bytes = 0;
  // profiling
  int profiling = 0;
  cudaEvent_t ev0[kProfSlots], ev1[kProfSlots];
  int prof_n = 0;
  bool prof_events = false;
  const char* last_algo = "none";
  bool attrs_set = false;
};

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<encode_tiled_fn>(p);
  return fn;
}

template <typename P>
static int grow(P** ptr, size_t* cur, size_t need) {
  if (need <= *cur) return 0;
  if (*ptr) {
    // a search enqueued earlier (on any stream of this device) may still use the old scratch: growth is
    // rare (sizes are stable after warm-up), so it simply waits for the device before freeing
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(*ptr);
  }
  *ptr = nullptr;
  *cur = 0;
  size_t bytes = std::max(need, static_cast<size_t>(256));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(ptr), bytes));
  *cur = bytes;
  return 0;
}

static size_t elem_bytes(const mips_index_s* h) { return h->dtype == MIPS_DTYPE_BF16 ? 2 : 4; }

static int encode_rows_tmap(CUtensorMap* out, void* base, int d_pad, int64_t rows, int box_rows) {
  encode_tiled_fn fn = get_encode_fn();
  if (!fn) return set_err(MIPS_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  const cuuint32_t estr[2] = {1, 1};
  const cuuint32_t box[2] = {tc::KCH, static_cast<cuuint32_t>(box_rows)};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(MIPS_E_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return 0;
}

static int encode_bank_tmap(mips_index_s* h) {
  h->tmap_valid = false;
  h->shadow_valid = false;
  if (h->d_pad > tc2::MAX_KCH * tc2::KCH) return 0;
  int rc;
  if (h->dtype == MIPS_DTYPE_BF16) {
    if ((rc = encode_rows_tmap(&h->tmap128, h->bank, h->d_pad, h->capacity, 128))) return rc;
    if ((rc = encode_rows_tmap(&h->tmap64, h->bank, h->d_pad, h->capacity, 64))) return rc;
    h->tmap_valid = true;
  } else if (h->shadow) {
    if ((rc = encode_rows_tmap(&h->tmap_shadow64, h->shadow, h->d_pad, h->capacity, 64))) return rc;
    h->shadow_valid = true;
  }
  return 0;
}

static int encode_query_tmap(mips_index_s* h, const void* q_bf16, int rows) {
  if (h->tmap_q_ptr == q_bf16 && h->tmap_q_rows == rows) return 0;
  int rc = encode_rows_tmap(&h->tmap_q, const_cast<void*>(q_bf16), h->d_pad, rows, tc2::BLOCK_M);
  if (rc) return rc;
  h->tmap_q_ptr = q_bf16;
  h->tmap_q_rows = rows;
  return 0;
}

// (re)allocate the shard to hold at least `rows`; preserves existing rows.
namespace {
struct DevBufGuard {   // frees what ensure_capacity allocated unless it is released to the index
  void* p[3] = {nullptr, nullptr, nullptr};
  ~DevBufGuard() {
    for (void* q : p)
      if (q) cudaFree(q);
  }
  void release() { p[0] = p[1] = p[2] = nullptr; }
};
}  // namespace

static int ensure_capacity(mips_index_s* h, int64_t rows, cudaStream_t st) {
  if (rows <= h->capacity) return 0;
  int64_t cap = std::max<int64_t>(rows, h->capacity * 2);
  cap = std::max<int64_t>(round_up_l(cap, kRowAlign), 1024);
  DevBufGuard g;
  const size_t row_bytes = static_cast<size_t>(h->d_pad) * elem_bytes(h);
  const size_t srow_bytes = static_cast<size_t>(h->d_pad) * 2;
  CUDA_TRY(cudaMalloc(&g.p[0], static_cast<size_t>(cap) * row_bytes));
  CUDA_TRY(cudaMalloc(&g.p[1], static_cast<size_t>(cap) * sizeof(float)));
  const bool want_shadow = h->dtype == MIPS_DTYPE_F32 && h->d_pad <= tc2::MAX_KCH * tc2::KCH;
  if (want_shadow) CUDA_TRY(cudaMalloc(&g.p[2], static_cast<size_t>(cap) * srow_bytes));
  void* nb = g.p[0];
  float* nn = static_cast<float*>(g.p[1]);
  __nv_bfloat16* ns = static_cast<__nv_bfloat16*>(g.p[2]);
  if (ns) {
    if (h->ntotal > 0 && h->shadow)
      CUDA_TRY(cudaMemcpyAsync(ns, h->shadow, static_cast<size_t>(h->ntotal) * srow_bytes, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemsetAsync(reinterpret_cast<uint8_t*>(ns) + static_cast<size_t>(h->ntotal) * srow_bytes, 0,
                             static_cast<size_t>(cap - h->ntotal) * srow_bytes, st));
  }
  if (h->ntotal > 0) {
    CUDA_TRY(cudaMemcpyAsync(nb, h->bank, static_cast<size_t>(h->ntotal) * row_bytes,
                             cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(nn, h->norm2, static_cast<size_t>(h->ntotal) * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  }
  // rows beyond ntotal must hold finite values: they are multiplied (and masked) by K1
  CUDA_TRY(cudaMemsetAsync(static_cast<uint8_t*>(nb) + static_cast<size_t>(h->ntotal) * row_bytes, 0,
                           static_cast<size_t>(cap - h->ntotal) * row_bytes, st));
  CUDA_TRY(cudaMemsetAsync(nn + h->ntotal, 0, static_cast<size_t>(cap - h->ntotal) * sizeof(float), st));
  // the old allocation may still be read by searches on other streams of this device: a reallocation
  // waits for all of them (rare: growth is geometric, or never with capacity_rows)
  CUDA_TRY(cudaDeviceSynchronize());
  g.release();
  if (h->bank) cudaFree(h->bank);
  if (h->norm2) cudaFree(h->norm2);
  if (h->shadow) cudaFree(h->shadow);
  h->shadow = ns;
  h->bank = nb;
  h->norm2 = nn;
  h->capacity = cap;
  return encode_bank_tmap(h);
}

static int set_kernel_attrs(mips_index_s* h) {
  if (h->attrs_set) return 0;
  const int simt_max = static_cast<int>(simt::smem_bytes(16));
  CUDA_TRY(cudaFuncSetAttribute(simt::search_simt_kernel<float, false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, simt_max));
  CUDA_TRY(cudaFuncSetAttribute(simt::search_simt_kernel<float, true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, simt_max));
  CUDA_TRY(cudaFuncSetAttribute(simt::search_simt_kernel<__nv_bfloat16, false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, simt_max));
  CUDA_TRY(cudaFuncSetAttribute(simt::search_simt_kernel<__nv_bfloat16, true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, simt_max));
#define TC_ATTR(L2, N, K)                                                                      \
  CUDA_TRY(cudaFuncSetAttribute(tc::search_tc_kernel<L2, N, K>,                                \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT))
  TC_ATTR(false, 128, 3); TC_ATTR(true, 128, 3); TC_ATTR(false, 128, 2); TC_ATTR(true, 128, 2);
  TC_ATTR(false, 64, 6);  TC_ATTR(true, 64, 6);  TC_ATTR(false, 64, 4);  TC_ATTR(true, 64, 4);
  TC_ATTR(false, 64, 3);  TC_ATTR(false, 64, 12); TC_ATTR(false, 64, 2);
#undef TC_ATTR
  CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6144 * 4 * 8));
  CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6144 * 4 * 8));
#define TC2_ATTR(L2, K)                                                                        \
  CUDA_TRY(cudaFuncSetAttribute(tc2::search_tc2_kernel<L2, K, false>,                          \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_LIMIT)); \
  CUDA_TRY(cudaFuncSetAttribute(tc2::search_tc2_kernel<L2, K, true>,                           \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_LIMIT))
  TC2_ATTR(false, 4); TC2_ATTR(true, 4); TC2_ATTR(false, 6); TC2_ATTR(true, 6);
  TC2_ATTR(false, 2); TC2_ATTR(true, 2);
#undef TC2_ATTR
  h->attrs_set = true;
  return 0;
}

// ------------------------------------------------------------------------------------------
extern "C" {

const char* mips_last_error(void) { return g_err; }
int64_t mips_launch_count(void) { return g_launches.load(); }

int mips_create(mips_handle* out, int d, int metric, int dtype, int device, int64_t capacity_rows) {
  if (!out) return set_err(MIPS_E_INVALID, "out is NULL");
  *out = nullptr;
  if (d <= 0 || d > 65536) return set_err(MIPS_E_INVALID, "d must be in [1, 65536], got %d", d);
  if (metric != MIPS_METRIC_IP && metric != MIPS_METRIC_L2)
    return set_err(MIPS_E_INVALID, "metric must be 0 (inner product) or 1 (L2), got %d", metric);
  if (dtype != MIPS_DTYPE_F32 && dtype != MIPS_DTYPE_BF16)
    return set_err(MIPS_E_INVALID, "dtype must be 0 (fp32) or 1 (bf16), got %d", dtype);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    return set_err(MIPS_E_INVALID, "device %d out of range (%d CUDA devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(MIPS_E_UNSUPPORTED, "built for sm_100a only; device %d is sm_%d%d", device,
                   prop.major, prop.minor);
  mips_index_s* h = new (std::nothrow) mips_index_s();
  if (!h) return set_err(MIPS_E_NOMEM, "host allocation failed");
  h->d = d;
  h->d_pad = round_up_i(d, kDimAlign);
  h->metric = metric;
  h->dtype = dtype;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->max_norm2_bits), 8 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->max_norm2_bits, 0, 8 * sizeof(unsigned int));
  h->fb_count_dev = reinterpret_cast<int*>(h->max_norm2_bits + 3);
  if (e != cudaSuccess) {
    delete h;
    return set_err(MIPS_E_CUDA, "cudaMalloc stats: %s", cudaGetErrorString(e));
  }
  int rc = set_kernel_attrs(h);
  if (rc == 0 && capacity_rows > 0) rc = ensure_capacity(h, capacity_rows, nullptr);
  if (rc != 0) {
    mips_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int mips_destroy(mips_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  void* dev[] = {h->bank, h->norm2, h->max_norm2_bits, h->q_prep, h->q_norm2, h->part_key,
                 h->part_ids, h->ign_local, h->after_row, h->hafter_key, h->hafter_id, h->pace, h->pool, h->shadow, h->q_hi, h->q_res2, h->cand_key, h->cand_rows, h->fb_flags, h->stage_x, h->sh_local, h->sh_gath, h->sh_qn2, h->dp_q, h->dp_ign, h->hq, h->hign, h->hkey, h->hids,
                 h->hxn2, h->hqn2, h->hD, h->hI};
  for (void* p : dev)
    if (p) cudaFree(p);
  if (h->prof_events)
    for (int i = 0; i < kProfSlots; ++i) {
      cudaEventDestroy(h->ev0[i]);
      cudaEventDestroy(h->ev1[i]);
    }
  delete h;
  return 0;
}

int mips_reset_async(mips_handle h, void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  h->ntotal = 0;
  h->phi = 0.f;
  // stream ordered: searches enqueued earlier on `stream` still see their certificate maxima / counters
  CUDA_TRY(cudaMemsetAsync(h->max_norm2_bits, 0, 8 * sizeof(unsigned int), static_cast<cudaStream_t>(stream)));
  return 0;
}

int mips_reset(mips_handle h) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  // no stream to order against: wait for whatever is still searching this index (any stream of the device)
  CUDA_TRY(cudaDeviceSynchronize());
  return mips_reset_async(h, nullptr);
}

int64_t mips_ntotal(mips_handle h) { return h ? h->ntotal : -1; }
int64_t mips_capacity(mips_handle h) { return h ? h->capacity : -1; }
int mips_dim(mips_handle h) { return h ? h->d : -1; }
int mips_metric(mips_handle h) { return h ? h->metric : -1; }
int mips_dtype(mips_handle h) { return h ? h->dtype : -1; }
int mips_set_phi(mips_handle h, float phi) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  h->phi = phi;
  return 0;
}
float mips_get_phi(mips_handle h) { return h ? h->phi : 0.f; }
const char* mips_last_algo(mips_handle h) { return h ? h->last_algo : "none"; }
int64_t mips_fallback_queries(mips_handle h, int reset) {
  if (!h) return -1;
  if (cudaSetDevice(h->device) != cudaSuccess) return -1;
  int v = 0;
  if (cudaMemcpy(&v, h->fb_count_dev, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (reset) cudaMemset(h->fb_count_dev, 0, sizeof(int));
  return v;
}

int mips_set_profiling(mips_handle h, int on) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (on && !h->prof_events) {
    for (int i = 0; i < kProfSlots; ++i) {
      CUDA_TRY(cudaEventCreate(&h->ev0[i]));
      CUDA_TRY(cudaEventCreate(&h->ev1[i]));
    }
    h->prof_events = true;
  }
  h->profiling = on;
  h->prof_n = 0;
  return 0;
}

// Sum of the K1 launch durations (ms) recorded since profiling was (re)enabled; *n_launches
// receives how many launches that covers. Sync on the recorded events.
float mips_k1_ms_total(mips_handle h) {
  if (!h || !h->prof_events || h->prof_n == 0) return -1.f;
  float total = 0.f;
  const int n = std::min(h->prof_n, kProfSlots);
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(h->ev1[i]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, h->ev0[i], h->ev1[i]) != cudaSuccess) return -1.f;
    total += ms;
  }
  return total;
}
int mips_prof_count(mips_handle h) { return h ? std::min(h->prof_n, kProfSlots) : 0; }

// ------------------------------------------------------------------------------------------ K0
static int launch_ingest(mips_index_s* h, const float* x_dev, int64_t n, int64_t n_pad, int normalize,
                         void* out, float* norm2_out, unsigned int* maxbits, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>((n_pad + 7) / 8);
  if (blocks == 0) return 0;
  static const bool k0_scalar = env_int("MIPS_K0_SCALAR", 0) != 0;
  // vector path: 16-byte loads need d % 4 == 0 and 16-byte aligned rows on both sides
  const bool vec_ok = h->d % 4 == 0 && h->d_pad <= 1024 && reinterpret_cast<uintptr_t>(x_dev) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(out) % 16 == 0 && !k0_scalar;
  if (vec_ok) {
#define K0_VEC(T, NV)                                                                                \
  ingest_rows_vec_kernel<T, NV><<<blocks, 256, 0, st>>>(x_dev, n, n_pad, h->d, h->d_pad, normalize,  \
                                                        static_cast<T*>(out), norm2_out, maxbits)
    const int nvl = (h->d_pad / 4 + 31) / 32;   // float4 groups per lane
    if (h->dtype == MIPS_DTYPE_BF16) {
      if (nvl <= 2) K0_VEC(__nv_bfloat16, 2); else if (nvl <= 4) K0_VEC(__nv_bfloat16, 4);
      else if (nvl <= 6) K0_VEC(__nv_bfloat16, 6); else K0_VEC(__nv_bfloat16, 8);
    } else {
      if (nvl <= 2) K0_VEC(float, 2); else if (nvl <= 4) K0_VEC(float, 4);
      else if (nvl <= 6) K0_VEC(float, 6); else K0_VEC(float, 8);
    }
#undef K0_VEC
    LAUNCH_CHECK("ingest_rows_vec_kernel");
    return 0;
  }
  if (h->dtype == MIPS_DTYPE_BF16)
    ingest_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        x_dev, n, n_pad, h->d, h->d_pad, normalize, static_cast<__nv_bfloat16*>(out), norm2_out, maxbits);
  else
    ingest_rows_kernel<float><<<blocks, 256, 0, st>>>(x_dev, n, n_pad, h->d, h->d_pad, normalize,
                                                      static_cast<float*>(out), norm2_out, maxbits);
  LAUNCH_CHECK("ingest_rows_kernel");
  return 0;
}

// fp32 index: bf16 shadow of the freshly stored rows [row0, row0 + n) + running maxima for the
// exactness certificate (k3_rerank.cuh)
static int update_shadow(mips_index_s* h, int64_t row0, int64_t n, cudaStream_t st) {
  if (h->dtype != MIPS_DTYPE_F32 || !h->shadow || n <= 0) return 0;
  shadow_rows_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, st>>>(
      static_cast<const float*>(h->bank) + static_cast<size_t>(row0) * h->d_pad, n, h->d_pad,
      h->shadow + static_cast<size_t>(row0) * h->d_pad, nullptr, h->max_norm2_bits + 1, h->max_norm2_bits + 2);
  LAUNCH_CHECK("shadow_rows_kernel");
  return 0;
}

int mips_add(mips_handle h, const float* x, int64_t n, int x_on_device, int normalize, void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (n < 0 || (n > 0 && !x)) return set_err(MIPS_E_INVALID, "bad x / n");
  if (n == 0) return 0;
  if (h->ntotal + n > (static_cast<int64_t>(1) << 31) - 64)
    return set_err(MIPS_E_UNSUPPORTED, "shard limited to 2^31-64 rows (int32 local ids)");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_capacity(h, h->ntotal + n, st);
  if (rc) return rc;
  const size_t eb = elem_bytes(h);
  if (x_on_device) {
    void* dst = static_cast<uint8_t*>(h->bank) + static_cast<size_t>(h->ntotal) * h->d_pad * eb;
    rc = launch_ingest(h, x, n, n, normalize, dst, h->norm2 + h->ntotal, h->max_norm2_bits, st);
    if (rc) return rc;
    if ((rc = update_shadow(h, h->ntotal, n, st))) return rc;
    h->ntotal += n;
    return 0;
  }
  // host rows: stage through a device buffer in <= 64 MiB chunks (stream ordered)
  const int64_t chunk_rows = std::max<int64_t>(1, (64ll << 20) / (static_cast<int64_t>(h->d) * 4));
  rc = grow(&h->stage_x, &h->stage_x_bytes,
            static_cast<size_t>(std::min(chunk_rows, n)) * h->d * sizeof(float));
  if (rc) return rc;
  for (int64_t r = 0; r < n; r += chunk_rows) {
    const int64_t m = std::min(chunk_rows, n - r);
    CUDA_TRY(cudaMemcpyAsync(h->stage_x, x + r * h->d, static_cast<size_t>(m) * h->d * sizeof(float),
                             cudaMemcpyHostToDevice, st));
    void* dst = static_cast<uint8_t*>(h->bank) + static_cast<size_t>(h->ntotal) * h->d_pad * eb;
    rc = launch_ingest(h, h->stage_x, m, m, normalize, dst, h->norm2 + h->ntotal, h->max_norm2_bits, st);
    if (rc) return rc;
    if ((rc = update_shadow(h, h->ntotal, m, st))) return rc;
    h->ntotal += m;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int mips_normalize_l2(float* x, int64_t n, int d, int x_on_device, int device, void* stream) {
  if (n < 0 || d <= 0 || (n > 0 && !x)) return set_err(MIPS_E_INVALID, "bad x / n / d");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* buf = x;
  const size_t bytes = static_cast<size_t>(n) * d * sizeof(float);
  if (!x_on_device) {
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&buf), bytes));
    cudaError_t e = cudaMemcpyAsync(buf, x, bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      cudaFree(buf);
      return set_err(MIPS_E_CUDA, "H2D: %s", cudaGetErrorString(e));
    }
  }
  // in place: each warp reads its whole row before it rewrites it element by element
  ingest_rows_kernel<float><<<static_cast<unsigned>((n + 7) / 8), 256, 0, st>>>(buf, n, n, d, d, 1, buf,
                                                                             nullptr, nullptr);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && !x_on_device) {
    e = cudaMemcpyAsync(x, buf, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  if (!x_on_device) cudaFree(buf);
  if (e != cudaSuccess) return set_err(MIPS_E_CUDA, "normalize_l2: %s", cudaGetErrorString(e));
  return 0;
}

int mips_max_norm2(mips_handle h, float* out, void* stream) {
  if (!h || !out) return set_err(MIPS_E_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int bits = 0;
  CUDA_TRY(cudaMemcpyAsync(&bits, h->max_norm2_bits, sizeof(bits), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(out, &bits, sizeof(float));
  return 0;
}

int mips_reconstruct(mips_handle h, int64_t row0, int64_t n, float* out, int out_on_device, void* stream) {
  if (!h || !out) return set_err(MIPS_E_INVALID, "null argument");
  if (row0 < 0 || n < 0 || row0 + n > h->ntotal)
    return set_err(MIPS_E_INVALID, "rows [%lld, %lld) outside [0, %lld)", (long long)row0,
                   (long long)(row0 + n), (long long)h->ntotal);
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunk_rows = out_on_device ? n : std::max<int64_t>(1, (64ll << 20) / (static_cast<int64_t>(h->d) * 4));
  if (!out_on_device) {
    int rc = grow(&h->stage_x, &h->stage_x_bytes,
                  static_cast<size_t>(std::min(chunk_rows, n)) * h->d * sizeof(float));
    if (rc) return rc;
  }
  for (int64_t r = 0; r < n; r += chunk_rows) {
    const int64_t m = std::min(chunk_rows, n - r);
    float* dst = out_on_device ? out + r * h->d : h->stage_x;
    const int64_t total = m * h->d;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    if (h->dtype == MIPS_DTYPE_BF16)
      reconstruct_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(h->bank), row0 + r, m, h->d, h->d_pad, dst);
    else
      reconstruct_rows_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(h->bank),
                                                             row0 + r, m, h->d, h->d_pad, dst);
    LAUNCH_CHECK("reconstruct_rows_kernel");
    if (!out_on_device) {
      CUDA_TRY(cudaMemcpyAsync(out + r * h->d, h->stage_x, static_cast<size_t>(total) * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ K1
// K1, CTA-pair tensor-core kernel over a bf16 row matrix described by `tmap_bank` (the bf16 bank, or
// the bf16 shadow of an fp32 bank). Leaves [n_splits, nq, k] candidate lists in h->part_key / part_ids.
static int launch_tc2(mips_index_s* h, const CUtensorMap& tmap_bank, const __nv_bfloat16* q_bf16, int nq,
                      int nq_pad, int k, const int* ign_local, bool l2, cudaStream_t st, int* n_parts_out,
                      bool allow_pool = true, const float* after_key = nullptr, const int* after_row = nullptr) {
    int rc = encode_query_tmap(h, q_bf16, nq_pad);
    if (rc) return rc;
    const int n_tiles = static_cast<int>((h->ntotal + tc2::TILE_N - 1) / tc2::TILE_N);
    const int n_qpairs = nq_pad / tc2::PAIR_M;
    const int n_splits = std::max(1, std::min((h->sm_count / 2) / n_qpairs, n_tiles));
    *n_parts_out = n_splits;
    const size_t pk = static_cast<size_t>(n_splits) * nq * k;
    rc = grow(&h->part_key, &h->part_key_bytes, pk * sizeof(float));
    if (rc) return rc;
    rc = grow(&h->part_ids, &h->part_ids_bytes, pk * sizeof(int));
    if (rc) return rc;
    tc2::Params p;
    p.q = q_bf16;
    p.xnorm2 = h->norm2;
    p.ignore_local = ign_local;
    p.after_key = after_key;
    p.after_row = after_row;
    p.part_key = h->part_key;
    p.part_ids = h->part_ids;
    p.ntotal = h->ntotal;
    p.nq = nq;
    p.d_pad = h->d_pad;
    p.k = k;
    p.n_tiles = n_tiles;
    p.n_qpairs = n_qpairs;
    p.n_splits = n_splits;
    // 48 KiB stages when three of them fit (k <= 8 at d = 768), else 32 KiB: A/B on one board with pacing on,
    // 12.16 ms vs 12.33 ms per launch on the 10M x 768 bank
    // (one query pair = HBM bound: 32 KiB stages, one more of them in flight: 10M x 768, nq=256: 3.05 vs 3.12 ms)
    int skch = (h->d_pad / tc2::KCH) % 6 == 0 && tc2::pick_stages(h->d_pad, k, 6) >= 3 && n_qpairs > 1 ? 6 : 4;
    static const int skch_env = env_int("MIPS_TC2_SKCH", 0);   // tuning experiments only
    if (skch_env == 2 || skch_env == 4 || skch_env == 6) skch = skch_env;
    while (skch > 2 && tc2::pick_stages(h->d_pad, k, skch) < 2) skch -= 2;
    p.stages = tc2::pick_stages(h->d_pad, k, skch);
    if (p.stages < 2)
      return set_err(MIPS_E_UNSUPPORTED, "tc2: d_pad=%d with k=%d does not fit shared memory", h->d_pad, k);
    p.cache_hint = n_qpairs > 1 ? ptx::kEvictNormal : ptx::kEvictFirst;
    p.pace = nullptr;
    p.pace_window = 6;   // measured: HBM reads 40 GB -> ~19 GB per launch on the 10M x 768 bank; step time within +-4 % of unpaced (which side wins depends on how hard the board is power-capped)
    static const int pace_env = env_int("MIPS_TC2_PACE", -1);   // tuning; 0 disables
    if (pace_env >= 0) p.pace_window = pace_env;
    // with two pairs per split the stalls cost more than the L2 hits save (nq=512: 7.03 ms unpaced vs 7.62 ms)
    if (n_qpairs >= 4 && p.pace_window > 0) {
      const size_t pb = static_cast<size_t>(n_splits) * n_qpairs * sizeof(int);
      rc = grow(&h->pace, &h->pace_bytes, pb);
      if (rc) return rc;
      CUDA_TRY(cudaMemsetAsync(h->pace, 0, pb, st));
      p.pace = h->pace;
    }
    // pooled admission threshold across the splits of a query (k1_topk.cuh). Measured on one board (pool on / off):
    // 250k rows x 1024 queries, k=32: 0.42 / 0.45 ms; 10M rows, k=64: 12.3 / 12.6 ms; but k=8: 12.65 / 12.15 ms
    // (10M x 1024), 4.2 / 3.4 ms (10M x 256, 74 splits to poll): the polling loads and the extra registers cost
    // more than the few admissions a small k has to save — a separate kernel instance, used for k >= 32 only
    p.pool = nullptr;
    p.pool_m = (k + n_splits - 1) / n_splits;
    static const int pool_env = env_int("MIPS_TC2_POOL", -1);   // tuning: 0 never, 1 whenever it is valid
    const bool pool_wanted = pool_env < 0 ? k >= 32 : pool_env != 0;
    if (allow_pool && pool_wanted && n_splits >= 2 && p.pool_m <= 4) {
      const size_t pb = static_cast<size_t>(nq_pad) * n_splits * sizeof(uint32_t);
      rc = grow(&h->pool, &h->pool_bytes, pb);
      if (rc) return rc;
      CUDA_TRY(cudaMemsetAsync(h->pool, 0, pb, st));
      p.pool = h->pool;
    }
    const size_t smem = tc2::smem_bytes(h->d_pad, k, p.stages, skch);
    const unsigned grid = static_cast<unsigned>(2 * n_qpairs * n_splits);
#define TC2_LAUNCH(K)                                                                                         \
  do {                                                                                                        \
    if (p.pool) {                                                                                             \
      if (l2) tc2::search_tc2_kernel<true, K, true><<<grid, tc2::THREADS, smem, st>>>(tmap_bank, h->tmap_q, p);   \
      else    tc2::search_tc2_kernel<false, K, true><<<grid, tc2::THREADS, smem, st>>>(tmap_bank, h->tmap_q, p);  \
    } else {                                                                                                  \
      if (l2) tc2::search_tc2_kernel<true, K, false><<<grid, tc2::THREADS, smem, st>>>(tmap_bank, h->tmap_q, p);  \
      else    tc2::search_tc2_kernel<false, K, false><<<grid, tc2::THREADS, smem, st>>>(tmap_bank, h->tmap_q, p); \
    }                                                                                                         \
  } while (0)
    if (skch == 6) TC2_LAUNCH(6); else if (skch == 4) TC2_LAUNCH(4); else TC2_LAUNCH(2);
#undef TC2_LAUNCH
    LAUNCH_CHECK("search_tc2_kernel");
    return 0;
}

// K1, 1-CTA tensor-core kernel over a bf16 row matrix described by `tmap` (boxes of acc_n rows): the bf16 bank,
// or the bf16 shadow of an fp32 bank. Leaves [n_splits, nq, k] candidate sets in h->part_key / part_ids.
static int launch_tc(mips_index_s* h, const CUtensorMap& tmap, const __nv_bfloat16* q_bf16, int nq, int nq_pad, int k,
                     int acc_n, const int* ign_local, bool l2, cudaStream_t st, int* n_parts_out,
                     const float* after_key = nullptr, const int* after_row = nullptr) {
    const int n_tiles = static_cast<int>((h->ntotal + acc_n - 1) / acc_n);
    const int n_qtiles = nq_pad / tc::BLOCK_M;
    int n_splits = std::max(1, std::min(h->sm_count / n_qtiles, n_tiles));
    *n_parts_out = n_splits;
    const size_t pk = static_cast<size_t>(n_splits) * nq * k;
    int rc = grow(&h->part_key, &h->part_key_bytes, pk * sizeof(float));
    if (rc) return rc;
    rc = grow(&h->part_ids, &h->part_ids_bytes, pk * sizeof(int));
    if (rc) return rc;
    tc::Params p;
    p.q = q_bf16;
    p.xnorm2 = h->norm2;
    p.ignore_local = ign_local;
    p.after_key = after_key;
    p.after_row = after_row;
    p.part_key = h->part_key;
    p.part_ids = h->part_ids;
    p.ntotal = h->ntotal;
    p.nq = nq;
    p.d_pad = h->d_pad;
    p.k = k;
    p.n_tiles = n_tiles;
    p.n_qtiles = n_qtiles;
    p.n_splits = n_splits;
    int skch = tc::pick_skch(h->d_pad, acc_n);
    static const int tc_skch_env = env_int("MIPS_TC_SKCH", 0);   // tuning experiments only (IP metric, 64-row accumulators)
    if (acc_n == 64 && !l2 && (tc_skch_env == 2 || tc_skch_env == 3 || tc_skch_env == 4 || tc_skch_env == 6 || tc_skch_env == 12))
      skch = tc_skch_env;
    p.stages = tc::pick_stages(k, skch, acc_n);
    // a bank tile is re-read by the other query tiles from L2; with one query tile it is dead
    p.cache_hint = n_qtiles > 1 ? ptx::kEvictNormal : ptx::kEvictFirst;
    const size_t smem = tc::smem_bytes(k, p.stages, skch, acc_n);
    const unsigned grid = static_cast<unsigned>(n_qtiles * n_splits);
#define TC_LAUNCH(N, K, MAP)                                                               \
  do {                                                                                     \
    if (l2) tc::search_tc_kernel<true, N, K><<<grid, tc::THREADS, smem, st>>>(MAP, p);     \
    else    tc::search_tc_kernel<false, N, K><<<grid, tc::THREADS, smem, st>>>(MAP, p);    \
  } while (0)
    if (acc_n == 128) {
      if (skch == 3) TC_LAUNCH(128, 3, tmap); else TC_LAUNCH(128, 2, tmap);
    } else {
      if (skch == 6) TC_LAUNCH(64, 6, tmap);
      else if (skch == 12) tc::search_tc_kernel<false, 64, 12><<<grid, tc::THREADS, smem, st>>>(tmap, p);
      else if (skch == 3) tc::search_tc_kernel<false, 64, 3><<<grid, tc::THREADS, smem, st>>>(tmap, p);
      else if (skch == 2) tc::search_tc_kernel<false, 64, 2><<<grid, tc::THREADS, smem, st>>>(tmap, p);
      else TC_LAUNCH(64, 4, tmap);
    }
#undef TC_LAUNCH
    LAUNCH_CHECK("search_tc_kernel");
    return 0;
}

// K1, exact fp32-FMA kernel over the stored rows (any dtype). `tile_active` (device, one int per
// 64-query tile, or null) restricts the launch to the query tiles that need recomputing.
static int launch_simt(mips_index_s* h, int nq, int k, const int* ign_local, bool l2, const int* tile_active,
                       cudaStream_t st, int* n_parts_out, const float* after_key = nullptr,
                       const int* after_row = nullptr) {
  const int n_tiles = static_cast<int>((h->ntotal + simt::BN - 1) / simt::BN);
  const int n_qtiles = (nq + simt::BM - 1) / simt::BM;
  const int target_blocks = 4 * h->sm_count;
  int n_splits = std::max(1, std::min((target_blocks + n_qtiles - 1) / n_qtiles, n_tiles));
  n_splits = std::min(n_splits, 65535);
  const int nsub = simt::nsub_for_k(k);
  const int n_parts = n_splits * nsub;
  *n_parts_out = n_parts;
  const size_t pk = static_cast<size_t>(n_parts) * nq * k;
  int rc = grow(&h->part_key, &h->part_key_bytes, pk * sizeof(float));
  if (rc) return rc;
  rc = grow(&h->part_ids, &h->part_ids_bytes, pk * sizeof(int));
  if (rc) return rc;
  const dim3 grid(n_qtiles, n_splits);
  const size_t smem = simt::smem_bytes(k);
#define SIMT_LAUNCH(T, L2)                                                                         \
  simt::search_simt_kernel<T, L2><<<grid, simt::THREADS, smem, st>>>(                              \
      static_cast<const T*>(h->q_prep), static_cast<const T*>(h->bank), h->norm2, nq, h->ntotal,   \
      h->d_pad, k, ign_local, n_tiles, h->part_key, h->part_ids, tile_active, after_key, after_row)
  if (h->dtype == MIPS_DTYPE_BF16) {
    if (l2) SIMT_LAUNCH(__nv_bfloat16, true); else SIMT_LAUNCH(__nv_bfloat16, false);
  } else {
    if (l2) SIMT_LAUNCH(float, true); else SIMT_LAUNCH(float, false);
  }
#undef SIMT_LAUNCH
  LAUNCH_CHECK("search_simt_kernel");
  return 0;
}

// K2 in LOCAL mode: split lists -> one list per query (global ids, |x|^2 gathered). Candidates are
// staged in shared memory when there are more than a handful per query.
static int launch_merge_local(mips_index_s* h, const float* part_key, const int* part_ids, const float* bank_xn2,
                              int n_parts, int nq, int k_in, int k_out, int64_t id_offset, float* out_key,
                              int64_t* out_ids, float* out_xn2, void* out_packed, const int* q_active,
                              cudaStream_t st, const char* what, const XchgOut* xo = nullptr) {
  const int C = n_parts * k_in;
  // every query's candidates are staged in shared memory (8 bytes each); fewer queries (warps) per block when
  // one query needs a large staging area
  int cap = 0, wpb = 4;
  if (C > 0 && C <= 6144) cap = C;
  else if (C > 6144 && C <= 12288) { cap = C; wpb = 2; }
  else if (C > 12288 && C <= 24576) { cap = C; wpb = 1; }
  const size_t smem = static_cast<size_t>(wpb) * cap * sizeof(uint2);
  static const int k2_cut = env_int("MIPS_K2_CUT", 512);   // tuning: staged sets larger than this are radix-cut first
  merge_topk_kernel<true><<<(nq + wpb - 1) / wpb, 32 * wpb, smem, st>>>(
      part_key, part_ids, nullptr, bank_xn2, n_parts, nq, k_in, k_out, id_offset, nullptr, h->metric, MIPS_OUT_IP,
      0.f, nullptr, out_key, out_ids, out_xn2, nullptr, nullptr, 1.f, 0.f, nullptr, 0, nullptr,
      static_cast<PackedCand*>(out_packed), q_active, cap, xo ? *xo : XchgOut{nullptr, nullptr, nullptr, 0u, 0},
      XchgIn{nullptr, 0u, 0, nullptr, 0ull}, k2_cut);
  LAUNCH_CHECK(what);
  return 0;
}

// candidates kept by the tensor-core filter of the exact search: enough head room between the k-th
// exact key and the kc-th approximate key for the certificate to hold on non-degenerate data
static int tcx_candidates(int k) { return std::min(2 * MIPS_MAX_K, std::max(4 * k, k + 16)); }
// entries each bank split keeps for the filter: a split rarely holds more than a few of the global
// candidates, and whatever it drops is covered by the certificate (its threshold enters T)
static int tcx_split_list(int k) { return std::max(k, 16); }

static int search_chunk(mips_index_s* h, const float* q, int nq, int k, int q_normalize,
                        const int64_t* ignore_ids, int64_t id_offset, int algo, float* out_key,
                        int64_t* out_ids, float* out_xnorm2, float* out_qnorm2, void* out_packed,
                        cudaStream_t st, const XchgOut* xo = nullptr, const float* after_key = nullptr,
                        const int64_t* after_id = nullptr) {
  int rc;
  const size_t eb = elem_bytes(h);
  const int nq_pad = round_up_i(nq, (algo == MIPS_ALGO_TC2 || algo == MIPS_ALGO_TCX) ? tc2::PAIR_M : tc::BLOCK_M);
  rc = grow(&h->q_prep, &h->q_prep_bytes, static_cast<size_t>(nq_pad) * h->d_pad * eb);
  if (rc) return rc;
  rc = grow(&h->q_norm2, &h->q_norm2_bytes, static_cast<size_t>(nq_pad) * sizeof(float));
  if (rc) return rc;
  // query preparation (_prepare_query, mips.py:368-375): normalise, cast, pad, |q|^2
  rc = launch_ingest(h, q, nq, nq_pad, q_normalize, h->q_prep, h->q_norm2, nullptr, st);
  if (rc) return rc;
  if (out_qnorm2)
    CUDA_TRY(cudaMemcpyAsync(out_qnorm2, h->q_norm2, static_cast<size_t>(nq) * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  const int* ign_local = nullptr;
  if (ignore_ids) {
    rc = grow(&h->ign_local, &h->ign_bytes, static_cast<size_t>(nq) * sizeof(int));
    if (rc) return rc;
    ignore_to_local_kernel<<<(nq + 255) / 256, 256, 0, st>>>(ignore_ids, nq, id_offset, h->ntotal,
                                                             h->ign_local);
    LAUNCH_CHECK("ignore_to_local_kernel");
    ign_local = h->ign_local;
  }

  const int* after_row = nullptr;
  if (after_key) {
    rc = grow(&h->after_row, &h->after_row_bytes, static_cast<size_t>(nq) * sizeof(int));
    if (rc) return rc;
    bound_to_local_kernel<<<(nq + 255) / 256, 256, 0, st>>>(after_id, nq, id_offset, h->after_row);
    LAUNCH_CHECK("bound_to_local_kernel");
    after_row = h->after_row;
  }

  const bool l2 = h->metric == MIPS_METRIC_L2;
  const bool use_tc = algo == MIPS_ALGO_TC || algo == MIPS_ALGO_TC128;
  int n_parts = 0;
  const int slot = h->prof_n % kProfSlots;
  if (h->profiling) CUDA_TRY(prof_record(h->ev0[slot], st));

  if (algo == MIPS_ALGO_TCX) {
    // exact fp32 search: tensor-core filter over the bf16 shadow, exact re-rank, certificate, and
    // the exact SIMT kernel for the query tiles that fail it (k3_rerank.cuh)
    const int kc = tcx_candidates(k);   // candidates re-ranked per query
    const int m = tcx_split_list(k);    // entries each split keeps (>= k: a split may hold the whole top-k)
    const int n_qt = (nq + simt::BM - 1) / simt::BM;
    if ((rc = grow(&h->q_hi, &h->q_hi_bytes, static_cast<size_t>(nq_pad) * h->d_pad * 2))) return rc;
    if ((rc = grow(&h->q_res2, &h->q_res2_bytes, static_cast<size_t>(nq_pad) * sizeof(float)))) return rc;
    if ((rc = grow(&h->cand_key, &h->cand_key_bytes, static_cast<size_t>(nq) * kc * sizeof(float)))) return rc;
    if ((rc = grow(&h->cand_rows, &h->cand_rows_bytes, static_cast<size_t>(nq) * kc * sizeof(int64_t)))) return rc;
    if ((rc = grow(&h->fb_flags, &h->fb_flags_bytes, static_cast<size_t>(nq + n_qt) * sizeof(int)))) return rc;
    int* need_fb = h->fb_flags;
    int* tile_flag = h->fb_flags + nq;
    CUDA_TRY(cudaMemsetAsync(tile_flag, 0, static_cast<size_t>(n_qt) * sizeof(int), st));
    shadow_rows_kernel<<<static_cast<unsigned>((nq_pad + 7) / 8), 256, 0, st>>>(
        static_cast<const float*>(h->q_prep), nq_pad, h->d_pad, h->q_hi, h->q_res2, nullptr, nullptr);
    LAUNCH_CHECK("shadow_rows_kernel<queries>");
    // small batches are HBM bound: the 1-CTA kernel streams the shadow as fast as the pair kernel and does
    // not pad the batch to 256 queries
    if (nq <= tc::BLOCK_M && h->d_pad <= tc::MAX_DPAD)
      rc = launch_tc(h, h->tmap_shadow64, h->q_hi, nq, round_up_i(nq, tc::BLOCK_M), m, 64, ign_local, l2, st, &n_parts);
    else
      rc = launch_tc2(h, h->tmap_shadow64, h->q_hi, nq, nq_pad, m, ign_local, l2, st, &n_parts,
                      /*allow_pool=*/false);   // the certificate reasons about every split's OWN list
    if (rc) return rc;
    rc = launch_merge_local(h, h->part_key, h->part_ids, nullptr, n_parts, nq, m, kc, 0, h->cand_key, h->cand_rows,
                            nullptr, nullptr, nullptr, st, "merge_topk_kernel<candidates>");
    if (rc) return rc;
#define RERANK(L2)                                                                                         \
  rerank_exact_kernel<L2><<<nq, 256, 0, st>>>(                                                             \
      static_cast<const float*>(h->q_prep), static_cast<const float*>(h->bank), h->norm2, h->d_pad,       \
      h->cand_key, h->cand_rows, nq, kc, k, h->part_key, h->part_ids, n_parts, m, h->q_norm2, h->q_res2,   \
      h->max_norm2_bits + 1, h->max_norm2_bits + 2, id_offset, out_key, out_ids, out_xnorm2,               \
      static_cast<PackedCand*>(out_packed), need_fb, tile_flag, h->fb_count_dev)
    if (l2) RERANK(true); else RERANK(false);
#undef RERANK
    LAUNCH_CHECK("rerank_exact_kernel");
    rc = launch_simt(h, nq, k, ign_local, l2, tile_flag, st, &n_parts);
    if (rc) return rc;
    if (h->profiling) {
      CUDA_TRY(prof_record(h->ev1[slot], st));
      h->prof_n++;
    }
    rc = launch_merge_local(h, h->part_key, h->part_ids, h->norm2, n_parts, nq, k, k, id_offset, out_key, out_ids,
                            out_xnorm2, out_packed, need_fb, st, "merge_topk_kernel<fallback>");
    if (rc) return rc;
    h->last_algo = "tcx";
    return 0;
  }
  if (algo == MIPS_ALGO_TC2) {
    rc = launch_tc2(h, h->tmap64, static_cast<const __nv_bfloat16*>(h->q_prep), nq, nq_pad, k, ign_local, l2, st, &n_parts,
                    true, after_key, after_row);
    if (rc) return rc;
    h->last_algo = "tc2";
  } else if (use_tc) {
    const int acc_n = algo == MIPS_ALGO_TC128 ? 128 : 64;
    rc = launch_tc(h, acc_n == 128 ? h->tmap128 : h->tmap64, static_cast<const __nv_bfloat16*>(h->q_prep), nq, nq_pad, k,
                   acc_n, ign_local, l2, st, &n_parts, after_key, after_row);
    if (rc) return rc;
    h->last_algo = acc_n == 128 ? "tc128" : "tc";
  } else {
    rc = launch_simt(h, nq, k, ign_local, l2, nullptr, st, &n_parts, after_key, after_row);
    if (rc) return rc;
    h->last_algo = "simt";
  }
  if (h->profiling) {
    CUDA_TRY(prof_record(h->ev1[slot], st));
    h->prof_n++;
  }

  // local k-way merge: split lists -> one list per query, global ids, |x|^2 gathered
  rc = launch_merge_local(h, h->part_key, h->part_ids, h->norm2, n_parts, nq, k, k, id_offset, out_key, out_ids,
                          out_xnorm2, out_packed, nullptr, st, "merge_topk_kernel<local>", xo);
  if (rc) return rc;
  return 0;
}

static int search_local_impl(mips_handle h, const float* q, int nq, int k, int q_normalize,
                             const int64_t* ignore_ids, int64_t id_offset, int algo, float* out_key,
                             int64_t* out_ids, float* out_xnorm2, float* out_qnorm2, void* out_packed,
                             void* stream, const XchgOut* xo = nullptr, const float* after_key = nullptr,
                             const int64_t* after_id = nullptr) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (nq < 0 || (nq > 0 && (!q || (!xo && !out_packed && (!out_key || !out_ids)))))
    return set_err(MIPS_E_INVALID, "bad q / outputs");
  if (k < 1 || k > MIPS_MAX_K) return set_err(MIPS_E_INVALID, "k must be in [1, %d], got %d", MIPS_MAX_K, k);
  if (nq == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tc_ok = h->dtype == MIPS_DTYPE_BF16 && h->d_pad <= tc::MAX_DPAD && h->tmap_valid;
  const bool tc2_ok = h->dtype == MIPS_DTYPE_BF16 && h->d_pad <= tc2::MAX_KCH * tc2::KCH && h->tmap_valid &&
                      tc2::pick_stages(h->d_pad, k, 2) >= 2;
  const bool tcx_ok = h->dtype == MIPS_DTYPE_F32 && h->shadow_valid &&
                      tc2::pick_stages(h->d_pad, tcx_split_list(k), 2) >= 2;
  if ((after_key == nullptr) != (after_id == nullptr)) return set_err(MIPS_E_INVALID, "after_key and after_id go together");
  if (algo == MIPS_ALGO_AUTO && h->dtype == MIPS_DTYPE_F32) {
    static const int auto_tcx = env_int("MIPS_AUTO_TCX", 1);
    // a bounded pass (multi-pass search, k > MIPS_MAX_K) compares EXACT keys: the fp32 FMA kernel, not the filter
    algo = (tcx_ok && auto_tcx && !after_key) ? MIPS_ALGO_TCX : MIPS_ALGO_SIMT;
  }
  if (algo == MIPS_ALGO_TCX && after_key)
    return set_err(MIPS_E_UNSUPPORTED, "bounded passes on an fp32 bank run on the exact fp32 kernel (algo auto / simt)");
  if (algo == MIPS_ALGO_TCX && !tcx_ok)
    return set_err(MIPS_E_UNSUPPORTED, "exact tensor-core search needs an fp32 bank with d_pad <= %d (and k small enough for shared memory)", tc2::MAX_KCH * tc2::KCH);
  if (algo == MIPS_ALGO_AUTO) {
    static const int auto_tc2 = env_int("MIPS_AUTO_TC2", 1);
    // the CTA pair pays off once both CTAs hold live queries; small batches are HBM bound on 1-CTA tiles
    // (nq=128 on 10M x 768: 2.3 / 2.6 / 3.3 ms at k = 8 / 32 / 64 vs 3.3 / 3.6 / 4.2 ms on the pair kernel)
    if (tc2_ok && (auto_tc2 && nq > tc::BLOCK_M || !tc_ok)) algo = MIPS_ALGO_TC2;
    else algo = tc_ok ? MIPS_ALGO_TC : MIPS_ALGO_SIMT;
  }
  if ((algo == MIPS_ALGO_TC || algo == MIPS_ALGO_TC128) && !tc_ok)
    return set_err(MIPS_E_UNSUPPORTED, "tensor-core search needs a bf16 bank with d_pad <= %d", tc::MAX_DPAD);
  if (algo == MIPS_ALGO_TC2 && !tc2_ok)
    return set_err(MIPS_E_UNSUPPORTED, "CTA-pair tensor-core search needs a bf16 bank with d_pad <= %d (and k small enough for shared memory)", tc2::MAX_KCH * tc2::KCH);
  if (algo != MIPS_ALGO_TC && algo != MIPS_ALGO_TC128 && algo != MIPS_ALGO_TC2 && algo != MIPS_ALGO_TCX && algo != MIPS_ALGO_SIMT)
    return set_err(MIPS_E_INVALID, "unknown algo %d", algo);
  if (xo && (algo == MIPS_ALGO_TCX || nq > (algo != MIPS_ALGO_SIMT ? h->sm_count * tc::BLOCK_M : 16384)))
    return set_err(MIPS_E_UNSUPPORTED, "peer-memory exchange: single-chunk bf16 / SIMT searches only");
  if (h->ntotal == 0) {
    // faiss semantics on an empty index: ids -1
    int rc0 = launch_merge_local(h, nullptr, nullptr, nullptr, 0, nq, k, k, 0, out_key, out_ids, out_xnorm2,
                                 out_packed, nullptr, st, "merge_topk_kernel<empty>", xo);
    if (rc0) return rc0;
    if (out_qnorm2) CUDA_TRY(cudaMemsetAsync(out_qnorm2, 0, static_cast<size_t>(nq) * sizeof(float), st));
    return 0;
  }
  // bound the scratch: chunks of queries (one kernel launch each)
  const int chunk = algo != MIPS_ALGO_SIMT ? h->sm_count * tc::BLOCK_M : 16384;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int m = std::min(chunk, nq - q0);
    int rc = search_chunk(h, q + static_cast<size_t>(q0) * h->d, m, k, q_normalize,
                          ignore_ids ? ignore_ids + q0 : nullptr, id_offset, algo,
                          out_key ? out_key + static_cast<size_t>(q0) * k : nullptr,
                          out_ids ? out_ids + static_cast<size_t>(q0) * k : nullptr,
                          out_xnorm2 ? out_xnorm2 + static_cast<size_t>(q0) * k : nullptr,
                          out_qnorm2 ? out_qnorm2 + q0 : nullptr,
                          out_packed ? static_cast<PackedCand*>(out_packed) + static_cast<size_t>(q0) * k : nullptr,
                          st, xo, after_key ? after_key + q0 : nullptr, after_id ? after_id + q0 : nullptr);
    if (rc) return rc;
  }
  return 0;
}

int mips_search_local(mips_handle h, const float* q, int nq, int k, int q_normalize,
                      const int64_t* ignore_ids, int64_t id_offset, int algo, float* out_key,
                      int64_t* out_ids, float* out_xnorm2, float* out_qnorm2, void* stream) {
  return search_local_impl(h, q, nq, k, q_normalize, ignore_ids, id_offset, algo, out_key, out_ids,
                           out_xnorm2, out_qnorm2, nullptr, stream);
}

int mips_search_local_after(mips_handle h, const float* q, int nq, int k, int q_normalize, const int64_t* ignore_ids,
                            int64_t id_offset, int algo, const float* after_key, const int64_t* after_id,
                            float* out_key, int64_t* out_ids, float* out_xnorm2, float* out_qnorm2, void* out_packed,
                            void* stream) {
  return search_local_impl(h, q, nq, k, q_normalize, ignore_ids, id_offset, algo, out_key, out_ids, out_xnorm2,
                           out_qnorm2, out_packed, stream, nullptr, after_key, after_id);
}

int mips_search_local_packed(mips_handle h, const float* q, int nq, int k, int q_normalize,
                             const int64_t* ignore_ids, int64_t id_offset, int algo, void* out_packed,
                             float* out_qnorm2, void* stream) {
  if (!out_packed && nq > 0) return set_err(MIPS_E_INVALID, "out_packed is NULL");
  return search_local_impl(h, q, nq, k, q_normalize, ignore_ids, id_offset, algo, nullptr, nullptr, nullptr,
                           out_qnorm2, out_packed, stream);
}

// ------------------------------------------------------------------------------------------ K2
// FINAL merge: candidates of a query staged in shared memory (16 bytes each: key, position, int64 id)
static int final_merge_staging(int C, int* cap, int* wpb) {
  *cap = 0;
  *wpb = 4;
  if (C > 0 && C <= 3072) *cap = C;
  else if (C > 3072 && C <= 12288) { *cap = C; *wpb = 1; }
  if (static_cast<size_t>(*wpb) * *cap * 16 > 48 * 1024)   // opt-in per device; this entry point has no handle
    CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6144 * 4 * 8));
  return 0;
}

static int merge_impl(const float* cand_key, const int64_t* cand_ids, const float* cand_xnorm2,
                      const void* cand_packed, int n_parts,
               int nq, int k_in, int k_out, int metric, int out_mode, float phi, const float* q_norm2,
               const int64_t* ignore_ids, float* D, int64_t* I, float* cosine, float* doc_prob,
               float beta, float beta_bias, float* memory_bias, int mem_len, void* stream) {
  if (nq < 0 || n_parts < 0 || k_in < 1 || k_out < 1 || k_out > MIPS_MAX_K)
    return set_err(MIPS_E_INVALID, "bad merge shape (n_parts=%d nq=%d k_in=%d k_out=%d)", n_parts, nq, k_in, k_out);
  if (nq == 0) return 0;
  if (!D || !I || (n_parts > 0 && !cand_packed && (!cand_key || !cand_ids)))
    return set_err(MIPS_E_INVALID, "null buffers");
  if (out_mode < MIPS_OUT_IP || out_mode > MIPS_OUT_AUGL2) return set_err(MIPS_E_INVALID, "bad out_mode %d", out_mode);
  const bool need_xn2 = out_mode == MIPS_OUT_L2 || metric == MIPS_METRIC_L2 || cosine || doc_prob || memory_bias;
  if (need_xn2 && n_parts > 0 && !cand_xnorm2 && !cand_packed) return set_err(MIPS_E_INVALID, "cand_xnorm2 required for L2 / cosine outputs");
  if ((out_mode != MIPS_OUT_IP || cosine || doc_prob || memory_bias) && !q_norm2)
    return set_err(MIPS_E_INVALID, "q_norm2 required for L2 / cosine outputs");
  if (memory_bias && mem_len < 1) return set_err(MIPS_E_INVALID, "mem_len must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cap, wpb;
  int rc_st = final_merge_staging(n_parts * k_in, &cap, &wpb);
  if (rc_st) return rc_st;
  merge_topk_kernel<false><<<(nq + wpb - 1) / wpb, 32 * wpb, static_cast<size_t>(wpb) * cap * 16, st>>>(
      cand_key, cand_ids, cand_xnorm2, nullptr, n_parts, nq, k_in, k_out, 0, ignore_ids, metric, out_mode,
      phi, q_norm2, D, I, nullptr, cosine, doc_prob, beta, beta_bias, memory_bias, mem_len,
      static_cast<const PackedCand*>(cand_packed), nullptr, nullptr, cap);
  LAUNCH_CHECK("merge_topk_kernel<final>");
  return 0;
}

int mips_merge(const float* cand_key, const int64_t* cand_ids, const float* cand_xnorm2, int n_parts,
               int nq, int k_in, int k_out, int metric, int out_mode, float phi, const float* q_norm2,
               const int64_t* ignore_ids, float* D, int64_t* I, float* cosine, float* doc_prob,
               float beta, float beta_bias, float* memory_bias, int mem_len, void* stream) {
  return merge_impl(cand_key, cand_ids, cand_xnorm2, nullptr, n_parts, nq, k_in, k_out, metric, out_mode, phi,
                    q_norm2, ignore_ids, D, I, cosine, doc_prob, beta, beta_bias, memory_bias, mem_len, stream);
}

int mips_merge_packed(const void* cand_packed, int n_parts, int nq, int k_in, int k_out, int metric,
                      int out_mode, float phi, const float* q_norm2, const int64_t* ignore_ids, float* D,
                      int64_t* I, float* cosine, float* doc_prob, float beta, float beta_bias,
                      float* memory_bias, int mem_len, void* stream) {
  if (!cand_packed && n_parts > 0 && nq > 0) return set_err(MIPS_E_INVALID, "cand_packed is NULL");
  return merge_impl(nullptr, nullptr, nullptr, cand_packed, n_parts, nq, k_in, k_out, metric, out_mode, phi,
                    q_norm2, ignore_ids, D, I, cosine, doc_prob, beta, beta_bias, memory_bias, mem_len, stream);
}

// ------------------------------------------------------------------------------------------ e2e
int mips_search_host(mips_handle h, const float* xq, int nq, int k, int q_normalize,
                     const int64_t* ignore_ids, int out_mode, float* D, int64_t* I, void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (nq < 0 || (nq > 0 && (!xq || !D || !I))) return set_err(MIPS_E_INVALID, "bad host buffers");
  if (k < 1 || k > MIPS_MAX_K_MULTIPASS)
    return set_err(MIPS_E_INVALID, "k must be in [1, %d], got %d", MIPS_MAX_K_MULTIPASS, k);
  if (nq == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kp_max = std::min(k, MIPS_MAX_K);                 // results per pass
  const size_t nk = static_cast<size_t>(nq) * kp_max;
  int rc = 0;
  if ((rc = grow(&h->hq, &h->hq_bytes, static_cast<size_t>(nq) * h->d * sizeof(float)))) return rc;
  if ((rc = grow(&h->hkey, &h->hkey_bytes, nk * sizeof(float)))) return rc;
  if ((rc = grow(&h->hids, &h->hids_bytes, nk * sizeof(int64_t)))) return rc;
  if ((rc = grow(&h->hxn2, &h->hxn2_bytes, nk * sizeof(float)))) return rc;
  if ((rc = grow(&h->hqn2, &h->hqn2_bytes, static_cast<size_t>(nq) * sizeof(float)))) return rc;
  if ((rc = grow(&h->hD, &h->hD_bytes, nk * sizeof(float)))) return rc;
  if ((rc = grow(&h->hI, &h->hI_bytes, nk * sizeof(int64_t)))) return rc;
  if (k > MIPS_MAX_K) {
    if ((rc = grow(&h->hafter_key, &h->hafter_key_bytes, static_cast<size_t>(nq) * sizeof(float)))) return rc;
    if ((rc = grow(&h->hafter_id, &h->hafter_id_bytes, static_cast<size_t>(nq) * sizeof(int64_t)))) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(h->hq, xq, static_cast<size_t>(nq) * h->d * sizeof(float), cudaMemcpyHostToDevice, st));
  const int64_t* ign_dev = nullptr;
  if (ignore_ids) {
    if ((rc = grow(&h->hign, &h->hign_bytes, static_cast<size_t>(nq) * sizeof(int64_t)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->hign, ignore_ids, static_cast<size_t>(nq) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    ign_dev = h->hign;
  }
  // k <= MIPS_MAX_K: one pass. Larger k (faiss accepts any k; the reference asks for k + 1 with an ignore list,
  // mips.py:383-386): passes of MIPS_MAX_K results, each restricted to the rows strictly after the last result
  // of the pass before in the order (key descending, id ascending) — exact, and a pass costs one search.
  // the bound compares keys for EQUALITY across passes: every pass must compute a row's key with the same
  // arithmetic. The tensor kernels do (same tiles, same accumulation order); an fp32 bank's first pass would be
  // the filter + re-rank path, whose keys differ in the last bits from the fp32 FMA kernel of the bounded passes
  const int algo = (k > MIPS_MAX_K && h->dtype == MIPS_DTYPE_F32) ? MIPS_ALGO_SIMT : MIPS_ALGO_AUTO;
  for (int done = 0; done < k; done += MIPS_MAX_K) {
    const int kp = std::min(MIPS_MAX_K, k - done);
    rc = mips_search_local_after(h, h->hq, nq, kp, q_normalize, ign_dev, 0, algo,
                                 done ? h->hafter_key : nullptr, done ? h->hafter_id : nullptr, h->hkey, h->hids,
                                 h->hxn2, h->hqn2, nullptr, st);
    if (rc) return rc;
    rc = mips_merge(h->hkey, h->hids, h->hxn2, 1, nq, kp, kp, h->metric, out_mode, h->phi, h->hqn2, nullptr,
                    h->hD, h->hI, nullptr, nullptr, 1.f, 0.f, nullptr, 0, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(D + done, static_cast<size_t>(k) * sizeof(float), h->hD, static_cast<size_t>(kp) * sizeof(float),
                               static_cast<size_t>(kp) * sizeof(float), nq, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpy2DAsync(I + done, static_cast<size_t>(k) * sizeof(int64_t), h->hI,
                               static_cast<size_t>(kp) * sizeof(int64_t), static_cast<size_t>(kp) * sizeof(int64_t), nq,
                               cudaMemcpyDeviceToHost, st));
    if (done + kp < k) {
      take_last_kernel<<<(nq + 255) / 256, 256, 0, st>>>(h->hkey, h->hids, nq, kp, h->hafter_key, h->hafter_id);
      LAUNCH_CHECK("take_last_kernel");
    }
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------------------------------ metrics
int mips_retriever_metrics(const int64_t* ids, int nq, int k, const int64_t* row_aid, int64_t n_rows,
                           const int64_t* query_aid, const float* counts, float* per_query, float* out3,
                           float* pred_out, void* stream) {
  if (nq < 0 || k < 1 || k > MIPS_MAX_K_MULTIPASS)
    return set_err(MIPS_E_INVALID, "bad nq / k (k must be in [1, %d])", MIPS_MAX_K_MULTIPASS);
  if (nq == 0) return 0;
  if (!ids || !row_aid || !query_aid || !counts || !per_query || !out3) return set_err(MIPS_E_INVALID, "null buffers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  retrieval_metrics_rows_kernel<<<(nq + 3) / 4, 128, 0, st>>>(ids, nq, k, row_aid, n_rows, query_aid, counts,
                                                             per_query, pred_out);
  LAUNCH_CHECK("retrieval_metrics_rows_kernel");
  retrieval_metrics_mean_kernel<<<1, 256, 0, st>>>(per_query, nq, out3);
  LAUNCH_CHECK("retrieval_metrics_mean_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ gather
int mips_gather_rows(mips_handle h, const int64_t* ids, int64_t n, int64_t id_offset, float* out, void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (n < 0 || (n > 0 && (!ids || !out))) return set_err(MIPS_E_INVALID, "bad ids / out");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((n + 7) / 8);
  if (h->dtype == MIPS_DTYPE_BF16)
    gather_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(h->bank), h->ntotal,
                                                             h->d, h->d_pad, ids, n, id_offset, out);
  else
    gather_rows_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(h->bank), h->ntotal, h->d, h->d_pad,
                                                     ids, n, id_offset, out);
  LAUNCH_CHECK("gather_rows_kernel");
  return 0;
}

int mips_gather_tokens(const int32_t* store_ids, const int32_t* store_len, int64_t n_rows, int L, const int64_t* ids,
                       int64_t n, int32_t pad_id, int32_t bos_id, int32_t eos_id, int64_t* input_ids,
                       int64_t* attention_mask, int64_t* memory_attention_mask, int64_t* global_attention_mask,
                       void* stream) {
  if (n < 0 || L < 1 || n_rows < 0) return set_err(MIPS_E_INVALID, "bad n / L / n_rows");
  if (n == 0) return 0;
  if (!store_ids || !store_len || !ids || !input_ids || !attention_mask) return set_err(MIPS_E_INVALID, "null buffers");
  if (n > 0x7fffffff) return set_err(MIPS_E_INVALID, "too many rows to gather in one call");
  gather_tokens_kernel<<<static_cast<unsigned>(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      store_ids, store_len, n_rows, L, ids, pad_id, bos_id, eos_id, input_ids, attention_mask, memory_attention_mask,
      global_attention_mask);
  LAUNCH_CHECK("gather_tokens_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ peer-memory exchange
int mips_xchg_alloc(int device, int64_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64 || bytes <= 0) return set_err(MIPS_E_INVALID, "bad arguments");
  CUDA_TRY(cudaSetDevice(device));
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  cudaIpcMemHandle_t hnd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hnd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_err(MIPS_E_CUDA, "exchange buffer: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &hnd, 64);
  *ptr = p;
  return 0;
}

int mips_xchg_open(int device, const void* handle64, void** ptr) {
  if (!ptr || !handle64) return set_err(MIPS_E_INVALID, "bad arguments");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int mips_xchg_close(int device, void* ptr) {
  if (!ptr) return 0;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int mips_xchg_free(int device, void* ptr) {
  if (!ptr) return 0;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaFree(ptr));
  return 0;
}

int mips_search_local_xchg(mips_handle h, const float* q, int nq, int k, int q_normalize, const int64_t* ignore_ids,
                           int64_t id_offset, int algo, void* const* peer_bufs, uint32_t* const* peer_flags,
                           int n_peers, uint32_t seq, float* out_qnorm2, void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (!peer_bufs || !peer_flags || n_peers < 1 || n_peers > 32) return set_err(MIPS_E_INVALID, "bad peer arrays");
  XchgOut xo{reinterpret_cast<PackedCand* const*>(peer_bufs), peer_flags,
             reinterpret_cast<unsigned int*>(h->max_norm2_bits + 4), seq, n_peers};
  return search_local_impl(h, q, nq, k, q_normalize, ignore_ids, id_offset, algo, nullptr, nullptr, nullptr,
                           out_qnorm2, nullptr, stream, &xo);
}

int mips_merge_xchg(const void* my_buf, const uint32_t* my_flags, int n_ranks, uint32_t seq, int nq, int k_in, int k_out,
                    int metric, int out_mode, float phi, const float* q_norm2, const int64_t* ignore_ids, float* D,
                    int64_t* I, float* cosine, float* doc_prob, float beta, float beta_bias, float* memory_bias,
                    int mem_len, void* stream) {
  if (!my_buf || !my_flags || n_ranks < 1 || n_ranks > 32) return set_err(MIPS_E_INVALID, "bad exchange buffer");
  if (nq < 0 || k_in < 1 || k_out < 1 || k_out > MIPS_MAX_K) return set_err(MIPS_E_INVALID, "bad merge shape");
  if (nq == 0) return 0;
  if (!D || !I) return set_err(MIPS_E_INVALID, "null buffers");
  if (out_mode < MIPS_OUT_IP || out_mode > MIPS_OUT_AUGL2) return set_err(MIPS_E_INVALID, "bad out_mode %d", out_mode);
  if ((out_mode != MIPS_OUT_IP || cosine || doc_prob || memory_bias) && !q_norm2)
    return set_err(MIPS_E_INVALID, "q_norm2 required for L2 / cosine outputs");
  if (memory_bias && mem_len < 1) return set_err(MIPS_E_INVALID, "mem_len must be >= 1");
  static const int xchg_timeout_s = env_int("MIPS_XCHG_TIMEOUT_S", 120);
  int cap, wpb;
  int rc_st = final_merge_staging(n_ranks * k_in, &cap, &wpb);
  if (rc_st) return rc_st;
  merge_topk_kernel<false><<<(nq + wpb - 1) / wpb, 32 * wpb, static_cast<size_t>(wpb) * cap * 16,
                             static_cast<cudaStream_t>(stream)>>>(
      nullptr, nullptr, nullptr, nullptr, n_ranks, nq, k_in, k_out, 0, ignore_ids, metric, out_mode, phi, q_norm2, D, I,
      nullptr, cosine, doc_prob, beta, beta_bias, memory_bias, mem_len, static_cast<const PackedCand*>(my_buf), nullptr,
      nullptr, cap, XchgOut{nullptr, nullptr, nullptr, 0u, 0},
      XchgIn{my_flags, seq, n_ranks, const_cast<uint32_t*>(my_flags) + MIPS_XCHG_TIMEOUT_WORD,
             static_cast<unsigned long long>(std::max(1, xchg_timeout_s)) * 1000000000ull});
  LAUNCH_CHECK("merge_topk_kernel<final, peer exchange>");
  return 0;
}

int mips_xchg_timeout_seq(int device, const uint32_t* my_flags, uint32_t* out_seq) {
  if (!my_flags || !out_seq) return set_err(MIPS_E_INVALID, "bad arguments");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaMemcpy(out_seq, my_flags + MIPS_XCHG_TIMEOUT_WORD, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------ K3 (NCCL)
#define NCCL_TRY(expr)                                                                              \
  do {                                                                                              \
    int _r = (expr);                                                                                \
    if (_r != 0) return set_err(MIPS_E_NCCL, "%s: %s", #expr, nc->GetErrorString(_r));              \
  } while (0)

int mips_nccl_version(void) {
  const nccl_dl::Api* nc = nccl_dl::api();
  int v = 0;
  if (!nc || nc->GetVersion(&v) != 0) return -1;
  return v;
}

int mips_nccl_unique_id(void* id128) {
  if (!id128) return set_err(MIPS_E_INVALID, "id128 is NULL");
  const nccl_dl::Api* nc = nccl_dl::api();
  if (!nc) return set_err(MIPS_E_NCCL, "%s", nccl_dl::why_unavailable());
  nccl_dl::UniqueId id;
  NCCL_TRY(nc->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int mips_nccl_comm_init(void** comm, int n_ranks, int rank, const void* id128, int device) {
  if (!comm || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return set_err(MIPS_E_INVALID, "bad communicator arguments");
  *comm = nullptr;
  const nccl_dl::Api* nc = nccl_dl::api();
  if (!nc) return set_err(MIPS_E_NCCL, "%s", nccl_dl::why_unavailable());
  CUDA_TRY(cudaSetDevice(device));
  nccl_dl::UniqueId id;
  memcpy(&id, id128, sizeof(id));
  nccl_dl::Comm c = nullptr;
  NCCL_TRY(nc->CommInitRank(&c, n_ranks, id, rank));
  *comm = c;
  return 0;
}

int mips_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  const nccl_dl::Api* nc = nccl_dl::api();
  if (!nc) return set_err(MIPS_E_NCCL, "%s", nccl_dl::why_unavailable());
  NCCL_TRY(nc->CommDestroy(comm));
  return 0;
}

int mips_allgather_topk(mips_handle h, void* nccl_comm, const void* local_packed, void* gathered_packed, int nq, int k,
                        void* stream) {
  if (!h || !nccl_comm) return set_err(MIPS_E_INVALID, "null handle / communicator");
  if (nq < 0 || k < 1 || (nq > 0 && (!local_packed || !gathered_packed))) return set_err(MIPS_E_INVALID, "bad buffers");
  if (nq == 0) return 0;
  const nccl_dl::Api* nc = nccl_dl::api();
  if (!nc) return set_err(MIPS_E_NCCL, "%s", nccl_dl::why_unavailable());
  CUDA_TRY(cudaSetDevice(h->device));
  NCCL_TRY(nc->AllGather(local_packed, gathered_packed, static_cast<size_t>(nq) * k * sizeof(PackedCand), nccl_dl::kUint8,
                         nccl_comm, static_cast<cudaStream_t>(stream)));
  return 0;
}

int mips_search_sharded(mips_handle h, void* nccl_comm, int n_ranks, const float* q, int nq, int k, int q_normalize,
                        const int64_t* ignore_ids, int64_t id_offset, int algo, int out_mode, float* D, int64_t* I,
                        float* cosine, float* doc_prob, float beta, float beta_bias, float* memory_bias, int mem_len,
                        void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (n_ranks < 1 || (n_ranks > 1 && !nccl_comm)) return set_err(MIPS_E_INVALID, "n_ranks > 1 needs a communicator");
  if (nq < 0 || (nq > 0 && (!q || !D || !I))) return set_err(MIPS_E_INVALID, "bad q / outputs");
  if (nq == 0) return 0;
  int rc;
  const size_t rec = static_cast<size_t>(nq) * k * sizeof(PackedCand);
  if ((rc = grow(&h->sh_local, &h->sh_local_bytes, rec))) return rc;
  if ((rc = grow(&h->sh_qn2, &h->sh_qn2_bytes, static_cast<size_t>(nq) * sizeof(float)))) return rc;
  if (n_ranks > 1 && (rc = grow(&h->sh_gath, &h->sh_gath_bytes, rec * n_ranks))) return rc;
  rc = mips_search_local_packed(h, q, nq, k, q_normalize, ignore_ids, id_offset, algo, h->sh_local, h->sh_qn2, stream);
  if (rc) return rc;
  const void* cand = h->sh_local;
  if (n_ranks > 1) {
    if ((rc = mips_allgather_topk(h, nccl_comm, h->sh_local, h->sh_gath, nq, k, stream))) return rc;
    cand = h->sh_gath;
  }
  return mips_merge_packed(cand, n_ranks, nq, k, k, h->metric, out_mode, h->phi, h->sh_qn2, nullptr, D, I, cosine,
                           doc_prob, beta, beta_bias, memory_bias, mem_len, stream);
}

int mips_search_sharded_dp(mips_handle h, void* nccl_comm, int n_ranks, int rank, const float* q_local, int nq_local,
                           int k, int q_normalize, const int64_t* ignore_local, int64_t id_offset, int algo,
                           int out_mode, float* D, int64_t* I, float* cosine, float* doc_prob, float beta,
                           float beta_bias, float* memory_bias, int mem_len, void* stream) {
  if (!h) return set_err(MIPS_E_INVALID, "null handle");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !nccl_comm))
    return set_err(MIPS_E_INVALID, "bad rank / n_ranks / communicator");
  if (nq_local < 0 || (nq_local > 0 && (!q_local || !D || !I))) return set_err(MIPS_E_INVALID, "bad q / outputs");
  if (nq_local == 0) return 0;   // every rank passes the same nq_local: nobody enters a collective
  if (n_ranks == 1)
    return mips_search_sharded(h, nullptr, 1, q_local, nq_local, k, q_normalize, ignore_local, id_offset, algo, out_mode,
                               D, I, cosine, doc_prob, beta, beta_bias, memory_bias, mem_len, stream);
  const nccl_dl::Api* nc = nccl_dl::api();
  if (!nc) return set_err(MIPS_E_NCCL, "%s", nccl_dl::why_unavailable());
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nq_all = nq_local * n_ranks;
  const size_t rec_local = static_cast<size_t>(nq_local) * k * sizeof(PackedCand);   // one rank's queries
  int rc;
  if ((rc = grow(&h->dp_q, &h->dp_q_bytes, static_cast<size_t>(nq_all) * h->d * sizeof(float)))) return rc;
  if ((rc = grow(&h->sh_local, &h->sh_local_bytes, rec_local * n_ranks))) return rc;
  if ((rc = grow(&h->sh_gath, &h->sh_gath_bytes, rec_local * n_ranks))) return rc;
  if ((rc = grow(&h->sh_qn2, &h->sh_qn2_bytes, static_cast<size_t>(nq_all) * sizeof(float)))) return rc;
  // 1. every rank learns every rank's queries (and ignored ids): [G * B, d], rank major
  NCCL_TRY(nc->AllGather(q_local, h->dp_q, static_cast<size_t>(nq_local) * h->d, nccl_dl::kFloat32, nccl_comm, st));
  const int64_t* ign_all = nullptr;
  if (ignore_local) {
    if ((rc = grow(&h->dp_ign, &h->dp_ign_bytes, static_cast<size_t>(nq_all) * sizeof(int64_t)))) return rc;
    NCCL_TRY(nc->AllGather(ignore_local, h->dp_ign, static_cast<size_t>(nq_local), nccl_dl::kInt64, nccl_comm, st));
    ign_all = h->dp_ign;
  }
  // 2. one local search of all G * B queries over this rank's shard
  rc = mips_search_local_packed(h, h->dp_q, nq_all, k, q_normalize, ign_all, id_offset, algo, h->sh_local, h->sh_qn2, stream);
  if (rc) return rc;
  // 3. all-to-all of the 16-byte records: rank r receives, from every shard, the lists of ITS B queries
  NCCL_TRY(nc->GroupStart());
  for (int g = 0; g < n_ranks; ++g) {
    int r1 = nc->Send(static_cast<const uint8_t*>(h->sh_local) + rec_local * g, rec_local, nccl_dl::kUint8, g, nccl_comm, st);
    int r2 = r1 ? r1 : nc->Recv(static_cast<uint8_t*>(h->sh_gath) + rec_local * g, rec_local, nccl_dl::kUint8, g, nccl_comm, st);
    if (r2) {
      nc->GroupEnd();
      return set_err(MIPS_E_NCCL, "ncclSend/ncclRecv: %s", nc->GetErrorString(r2));
    }
  }
  NCCL_TRY(nc->GroupEnd());
  // 4. each rank merges the G lists of its own queries (+ the fused doc-score outputs)
  return mips_merge_packed(h->sh_gath, n_ranks, nq_local, k, k, h->metric, out_mode, h->phi,
                           h->sh_qn2 + static_cast<size_t>(rank) * nq_local, nullptr, D, I, cosine, doc_prob, beta,
                           beta_bias, memory_bias, mem_len, stream);
}
#undef NCCL_TRY

// ------------------------------------------------------------------------------------------ mixture
static int mixture_device(const float* logits) {
  // the shared-memory opt-in is per DEVICE (and these entry points have no handle): run on the device that owns
  // `logits` and set the attributes there on every call (a few hundred ns on the host, no device work)
  cudaPointerAttributes pa;
  CUDA_TRY(cudaPointerGetAttributes(&pa, logits));
  if (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged)
    return set_err(MIPS_E_INVALID, "logits must be device memory");
  CUDA_TRY(cudaSetDevice(pa.device));
  CUDA_TRY(cudaFuncSetAttribute(mix::copy_mixture_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (mix::MAX_HALF + 4) * static_cast<int>(sizeof(float))));
  CUDA_TRY(cudaFuncSetAttribute(mix::copy_mixture_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (mix::MAX_HALF + 4) * static_cast<int>(sizeof(float))));
  return 0;
}

int mips_copy_mixture_fwd(const float* logits, const float* gen_gate, const float* copy_probs, const int64_t* copy_seq,
                          int64_t n_rows, int rows_per_batch, int V, int S, float eps, float* out, float* stats,
                          void* stream) {
  if (n_rows < 0 || rows_per_batch < 1 || V < 1 || S < 0) return set_err(MIPS_E_INVALID, "bad shape");
  if (n_rows == 0) return 0;
  if (n_rows % rows_per_batch != 0) return set_err(MIPS_E_INVALID, "n_rows must be a multiple of rows_per_batch");
  if (!logits || !gen_gate || !out || (S > 0 && (!copy_probs || !copy_seq))) return set_err(MIPS_E_INVALID, "null buffers");
  if (V > mix::MAX_V)
    return set_err(MIPS_E_UNSUPPORTED, "vocabulary of %d does not fit one CTA's shared memory (max %d)", V, mix::MAX_V);
  if (n_rows > 0x3fffffff) return set_err(MIPS_E_INVALID, "too many rows");
  int rc = mixture_device(logits);
  if (rc) return rc;
  mix::copy_mixture_kernel<<<static_cast<unsigned>(2 * n_rows), mix::THREADS, static_cast<size_t>((V + 1) / 2 + 4) * sizeof(float),
                             static_cast<cudaStream_t>(stream)>>>(logits, gen_gate, copy_probs, copy_seq,
                                                                  rows_per_batch, V, S, eps, out, stats);
  LAUNCH_CHECK("copy_mixture_kernel");
  return 0;
}

int mips_copy_mixture(const float* logits, const float* gen_gate, const float* copy_probs, const int64_t* copy_seq,
                      int64_t n_rows, int rows_per_batch, int V, int S, float eps, float* out, void* stream) {
  return mips_copy_mixture_fwd(logits, gen_gate, copy_probs, copy_seq, n_rows, rows_per_batch, V, S, eps, out, nullptr,
                               stream);
}

int mips_copy_mixture_bwd(const float* logits, const float* out, const float* dout, const float* gen_gate,
                          const float* stats, const int64_t* copy_seq, int64_t n_rows, int rows_per_batch, int V, int S,
                          float* dlogits, float* dgate, float* dcopy, void* stream) {
  if (n_rows < 0 || rows_per_batch < 1 || V < 1 || S < 0) return set_err(MIPS_E_INVALID, "bad shape");
  if (n_rows == 0) return 0;
  if (n_rows % rows_per_batch != 0) return set_err(MIPS_E_INVALID, "n_rows must be a multiple of rows_per_batch");
  if (!logits || !out || !dout || !gen_gate || !stats || !dlogits || !dgate || (S > 0 && (!copy_seq || !dcopy)))
    return set_err(MIPS_E_INVALID, "null buffers");
  if (V > mix::MAX_V)
    return set_err(MIPS_E_UNSUPPORTED, "vocabulary of %d does not fit one CTA's shared memory (max %d)", V, mix::MAX_V);
  if (n_rows > 0x3fffffff) return set_err(MIPS_E_INVALID, "too many rows");
  int rc = mixture_device(logits);
  if (rc) return rc;
  mix::copy_mixture_bwd_kernel<<<static_cast<unsigned>(2 * n_rows), mix::THREADS,
                                 static_cast<size_t>((V + 1) / 2 + 4) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      logits, out, dout, gen_gate, stats, copy_seq, rows_per_batch, V, S, dlogits, dgate, dcopy);
  LAUNCH_CHECK("copy_mixture_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ copy attention
int mips_copy_attention_softmax_fwd(const float* scores, const float* doc_scores, int n_docs, int mem_len, float beta,
                                    float beta_bias, const float* beta_dev, const float* mask, int B, int T, int S,
                                    float* probs, void* stream) {
  if (B < 0 || T < 0 || S < 1) return set_err(MIPS_E_INVALID, "bad shape");
  if (B == 0 || T == 0) return 0;
  if (!scores || !probs) return set_err(MIPS_E_INVALID, "null buffers");
  if (doc_scores && (n_docs < 1 || mem_len < 1 || static_cast<int64_t>(n_docs) * mem_len < S))
    return set_err(MIPS_E_INVALID, "doc_scores needs n_docs * mem_len >= S");
  if (S > cattn::MAX_S) return set_err(MIPS_E_UNSUPPORTED, "S = %d memory tokens exceed one CTA's shared memory (max %d)", S, cattn::MAX_S);
  const size_t smem = static_cast<size_t>(S) * sizeof(float);
  CUDA_TRY(cudaFuncSetAttribute(cattn::biased_softmax_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                cattn::MAX_S * static_cast<int>(sizeof(float))));
  cattn::biased_softmax_fwd_kernel<<<static_cast<unsigned>(static_cast<int64_t>(B) * T), cattn::THREADS, smem,
                                     static_cast<cudaStream_t>(stream)>>>(scores, doc_scores, n_docs, mem_len, beta, beta_bias,
                                                                          beta_dev, mask, T, S, probs);
  LAUNCH_CHECK("biased_softmax_fwd_kernel");
  return 0;
}

int mips_copy_attention_softmax_bwd(const float* probs, const float* dprobs, int n_docs, int mem_len, int B, int T, int S,
                                    float* dscores, float* doc_grad, void* stream) {
  if (B < 0 || T < 0 || S < 1) return set_err(MIPS_E_INVALID, "bad shape");
  if (B == 0 || T == 0) return 0;
  if (!probs || !dprobs || !dscores) return set_err(MIPS_E_INVALID, "null buffers");
  if (doc_grad && (n_docs < 1 || mem_len < 1 || static_cast<int64_t>(n_docs) * mem_len < S))
    return set_err(MIPS_E_INVALID, "doc_grad needs n_docs * mem_len >= S");
  if (doc_grad && n_docs > cattn::MAX_S)
    return set_err(MIPS_E_UNSUPPORTED, "%d documents exceed the per-row shared-memory bins (max %d)", n_docs, cattn::MAX_S);
  const size_t smem = doc_grad ? static_cast<size_t>(n_docs) * sizeof(float) : 0;
  CUDA_TRY(cudaFuncSetAttribute(cattn::biased_softmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                cattn::MAX_S * static_cast<int>(sizeof(float))));
  cattn::biased_softmax_bwd_kernel<<<static_cast<unsigned>(static_cast<int64_t>(B) * T), cattn::THREADS, smem,
                                     static_cast<cudaStream_t>(stream)>>>(probs, dprobs, n_docs, mem_len, T, S, dscores,
                                                                          doc_grad);
  LAUNCH_CHECK("biased_softmax_bwd_kernel");
  return 0;
}

}  // extern "C"
