// nccl_dl.h — NCCL reached through dlopen (K3 of SURVEY §8b: the cross-GPU step of the sharded search on the
// caller's stream, through the C ABI, no torch in the step). The library binds to the libnccl.so.2 that is
// ALREADY loaded in the process (torch's bundled 2.28 when the host is Python) and only otherwise loads the
// system one (a plain C host), so there is no link-time dependency and never two NCCL copies in one process.
// ncclComm_t crosses the ABI as void*. Only the handful of entry points the search path uses are bound; the
// prototypes below restate nccl.h (2.27/2.28: identical for these symbols).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
#include <cstdlib>
#include <mutex>

namespace nccl_dl {

struct UniqueId {
  char internal[128];   // NCCL_UNIQUE_ID_BYTES
};
typedef void* Comm;
enum DataType { kInt8 = 0, kUint8 = 1, kInt32 = 2, kInt64 = 4, kFloat32 = 7 };
enum RedOp { kSum = 0, kMax = 2 };

struct Api {
  void* so = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*CommAbort)(Comm) = nullptr;
  int (*CommGetAsyncError)(Comm, int*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

inline const Api* api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("MIPS_NCCL_LIB");
    if (env && *env) a.so = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!a.so) a.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already has
    if (!a.so) a.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.so) a.so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.so) return;
    bool ok = true;
#define NCCL_DL_BIND(field, sym)                                          \
  do {                                                                    \
    *reinterpret_cast<void**>(&a.field) = dlsym(a.so, sym);               \
    if (!a.field) ok = false;                                             \
  } while (0)
    NCCL_DL_BIND(GetVersion, "ncclGetVersion");
    NCCL_DL_BIND(GetUniqueId, "ncclGetUniqueId");
    NCCL_DL_BIND(CommInitRank, "ncclCommInitRank");
    NCCL_DL_BIND(CommDestroy, "ncclCommDestroy");
    NCCL_DL_BIND(CommAbort, "ncclCommAbort");
    NCCL_DL_BIND(CommGetAsyncError, "ncclCommGetAsyncError");
    NCCL_DL_BIND(AllGather, "ncclAllGather");
    NCCL_DL_BIND(AllReduce, "ncclAllReduce");
    NCCL_DL_BIND(Send, "ncclSend");
    NCCL_DL_BIND(Recv, "ncclRecv");
    NCCL_DL_BIND(GroupStart, "ncclGroupStart");
    NCCL_DL_BIND(GroupEnd, "ncclGroupEnd");
    NCCL_DL_BIND(GetErrorString, "ncclGetErrorString");
#undef NCCL_DL_BIND
    if (!ok) a.so = nullptr;
  });
  return a.so ? &a : nullptr;
}
inline const char* why_unavailable() {
  return api() ? "" : "NCCL unavailable: libnccl.so.2 could not be loaded or lacks a required symbol (set MIPS_NCCL_LIB)";
}

}  // namespace nccl_dl
