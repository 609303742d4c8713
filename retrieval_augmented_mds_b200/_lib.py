"""ctypes binding of libmips_b200.so (the C ABI in include/mips_b200.h).

There is no fallback: if the shared library is missing it is built with nvcc; if that fails, or
if a compute entry point is called without a B200, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

from . import build as _build

_HEADER = Path(__file__).resolve().parent.parent / "include" / "mips_b200.h"

METRIC_IP, METRIC_L2 = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
ALGO_AUTO, ALGO_SIMT, ALGO_TC, ALGO_TC128, ALGO_TC2, ALGO_TCX = 0, 1, 2, 3, 4, 5
OUT_IP, OUT_L2, OUT_AUGL2 = 0, 1, 2
MAX_K = 64                 # per-pass capacity of the search kernels (MIPS_MAX_K)
MAX_K_MULTIPASS = 2048     # largest k of the multi-pass paths (MIPS_MAX_K_MULTIPASS)

_lib = None


class MipsError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function include/mips_b200.h declares (used by the symbol-export test)."""
    text = re.sub(r"/\*.*?\*/", "", _HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mips_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("MIPS_B200_LIB")   # developer A/B runs: another build of the same ABI
    path = Path(override) if override else _build.build()
    L = C.CDLL(str(path))
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        "mips_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, i64]),
        "mips_destroy": (i32, [vp]),
        "mips_reset": (i32, [vp]),
        "mips_reset_async": (i32, [vp, vp]),
        "mips_ntotal": (i64, [vp]),
        "mips_capacity": (i64, [vp]),
        "mips_dim": (i32, [vp]),
        "mips_metric": (i32, [vp]),
        "mips_dtype": (i32, [vp]),
        "mips_add": (i32, [vp, vp, i64, i32, i32, vp]),
        "mips_max_norm2": (i32, [vp, C.POINTER(f32), vp]),
        "mips_set_phi": (i32, [vp, f32]),
        "mips_get_phi": (f32, [vp]),
        "mips_normalize_l2": (i32, [vp, i64, i32, i32, i32, vp]),
        "mips_reconstruct": (i32, [vp, i64, i64, vp, i32, vp]),
        "mips_search_local": (i32, [vp, vp, i32, i32, i32, vp, i64, i32, vp, vp, vp, vp, vp]),
        "mips_search_local_after": (i32, [vp, vp, i32, i32, i32, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mips_merge": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp,
                             f32, f32, vp, i32, vp]),
        "mips_search_local_packed": (i32, [vp, vp, i32, i32, i32, vp, i64, i32, vp, vp, vp]),
        "mips_merge_packed": (i32, [vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, f32, f32,
                                    vp, i32, vp]),
        "mips_search_host": (i32, [vp, vp, i32, i32, i32, vp, i32, vp, vp, vp]),
        "mips_last_error": (C.c_char_p, []),
        "mips_launch_count": (i64, []),
        "mips_last_algo": (C.c_char_p, [vp]),
        "mips_fallback_queries": (i64, [vp, i32]),
        "mips_xchg_alloc": (i32, [i32, i64, C.POINTER(vp), vp]),
        "mips_xchg_open": (i32, [i32, vp, C.POINTER(vp)]),
        "mips_xchg_close": (i32, [i32, vp]),
        "mips_xchg_free": (i32, [i32, vp]),
        "mips_search_local_xchg": (i32, [vp, vp, i32, i32, i32, vp, i64, i32, vp, vp, i32, C.c_uint32, vp, vp]),
        "mips_merge_xchg": (i32, [vp, vp, i32, C.c_uint32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, f32,
                                  f32, vp, i32, vp]),
        "mips_xchg_timeout_seq": (i32, [i32, vp, C.POINTER(C.c_uint32)]),
        "mips_nccl_version": (i32, []),
        "mips_nccl_unique_id": (i32, [vp]),
        "mips_nccl_comm_init": (i32, [C.POINTER(vp), i32, i32, vp, i32]),
        "mips_nccl_comm_destroy": (i32, [vp]),
        "mips_allgather_topk": (i32, [vp, vp, vp, vp, i32, i32, vp]),
        "mips_search_sharded": (i32, [vp, vp, i32, vp, i32, i32, i32, vp, i64, i32, i32, vp, vp, vp, vp, f32, f32, vp,
                                      i32, vp]),
        "mips_search_sharded_dp": (i32, [vp, vp, i32, i32, vp, i32, i32, i32, vp, i64, i32, i32, vp, vp, vp, vp, f32,
                                         f32, vp, i32, vp]),
        "mips_gather_rows": (i32, [vp, vp, i64, i64, vp, vp]),
        "mips_gather_tokens": (i32, [vp, vp, i64, i32, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp]),
        "mips_copy_mixture": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, f32, vp, vp]),
        "mips_copy_mixture_fwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, f32, vp, vp, vp]),
        "mips_copy_mixture_bwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp]),
        "mips_copy_attention_softmax_fwd": (i32, [vp, vp, i32, i32, f32, f32, vp, vp, i32, i32, i32, vp, vp]),
        "mips_copy_attention_softmax_bwd": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
        "mips_retriever_metrics": (i32, [vp, i32, i32, vp, i64, vp, vp, vp, vp, vp, vp]),
        "mips_set_profiling": (i32, [vp, i32]),
        "mips_k1_ms_total": (f32, [vp]),
        "mips_prof_count": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().mips_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise MipsError(f"libmips_b200 error {rc}: {msg}")
