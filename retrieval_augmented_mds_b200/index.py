"""B200FlatIndex — faiss-style flat exact index whose bank lives in HBM and whose search runs as
hand-written sm_100a kernels behind the C ABI of include/mips_b200.h.

It mirrors the part of the faiss `Index` protocol the reference consumes through HF `datasets`
(sotasum/mips.py:333-345 add_faiss_index / .nprobe, :383-386 faiss_index.search;
retriever_lightning.py:317-321, :400-404; pretrain.py:475-479, :519-523):
`d`, `ntotal`, `metric_type`, `is_trained`, `verbose`, `train`, `add`, `search`, `reset`.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib
from ._lib import (ALGO_AUTO, ALGO_SIMT, ALGO_TC, ALGO_TC128, ALGO_TC2, ALGO_TCX, DTYPE_BF16, DTYPE_F32, MAX_K, MAX_K_MULTIPASS,
                   OUT_AUGL2, OUT_IP, OUT_L2, check)

METRIC_INNER_PRODUCT = 0  # faiss.METRIC_INNER_PRODUCT
METRIC_L2 = 1  # faiss.METRIC_L2

_DTYPES = {"fp32": DTYPE_F32, "float32": DTYPE_F32, "f32": DTYPE_F32, torch.float32: DTYPE_F32,
           "bf16": DTYPE_BF16, "bfloat16": DTYPE_BF16, torch.bfloat16: DTYPE_BF16}
_ALGOS = {"auto": ALGO_AUTO, "simt": ALGO_SIMT, "tc": ALGO_TC, "tc128": ALGO_TC128, "tc2": ALGO_TC2, "tcx": ALGO_TCX}


def _device_index(device) -> int:
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("B200FlatIndex needs a CUDA device (sm_100a); there is no CPU path")
        return torch.cuda.current_device()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError(f"B200FlatIndex lives on a CUDA device, got {dev}")
    return dev.index if dev.index is not None else torch.cuda.current_device()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class B200FlatIndex:
    """Exact flat index (inner product or L2) over a bf16 or fp32 bank resident in HBM."""

    is_trained = True

    def __init__(self, d: int, metric_type: int = METRIC_INNER_PRODUCT, dtype="bf16", device=None,
                 capacity: int = 0, id_offset: int = 0):
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of 'bf16', 'fp32', got {dtype!r}")
        self._L = _lib.lib()
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        self._h = C.c_void_p()
        check(self._L.mips_create(C.byref(self._h), int(d), int(metric_type), _DTYPES[dtype],
                                  self.device_index, int(capacity)))
        self.d = int(d)
        self.metric_type = int(metric_type)
        self.dtype = "bf16" if _DTYPES[dtype] == DTYPE_BF16 else "fp32"
        self.id_offset = int(id_offset)  # global id of local row 0 (row-sharded banks)
        self.verbose = False
        self.nprobe = 1  # accepted and ignored: exact search (reference sets it at mips.py:342-345)

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.mips_destroy(h)
            h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter teardown
            pass

    @property
    def ntotal(self) -> int:
        return int(self._L.mips_ntotal(self._h))

    @property
    def capacity(self) -> int:
        return int(self._L.mips_capacity(self._h))

    def reset(self) -> None:
        """faiss Index.reset(), ordered on the current torch stream of the index's device."""
        with torch.cuda.device(self.device):
            check(self._L.mips_reset_async(self._h, self._stream()))

    def train(self, x=None) -> None:  # flat index: nothing to train (datasets calls it when train_size is set)
        return None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ build side
    def add(self, x, normalize: bool = False) -> None:
        """faiss Index.add: x float32 [n, d], numpy (host) or torch tensor (host or this device).
        normalize=True fuses Mips._map_normalize (mips.py:358-361) into the ingest."""
        if isinstance(x, torch.Tensor):
            if x.dim() != 2 or x.shape[1] != self.d:
                raise ValueError(f"Shape of vectors must be [n, {self.d}], got {tuple(x.shape)}")
            if x.is_cuda:
                if x.device != self.device:
                    raise ValueError(f"vectors on {x.device}, index on {self.device}")
                xc = x.detach().to(torch.float32).contiguous()
                with torch.cuda.device(self.device):
                    check(self._L.mips_add(self._h, _ptr(xc), xc.shape[0], 1, int(normalize), self._stream()))
                    # xc must outlive the enqueued kernel
                    xc.record_stream(torch.cuda.current_stream(self.device))
                return
            x = x.detach().to(torch.float32).contiguous().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"Shape of vectors must be [n, {self.d}], got {x.shape}")
        with torch.cuda.device(self.device):
            check(self._L.mips_add(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], 0, int(normalize),
                                   self._stream()))

    def max_norm2(self) -> float:
        """max_i |x_i|^2 of the rows as added = get_phi (mips.py:55-56) = max_norm**2 (:298-304)."""
        out = C.c_float()
        check(self._L.mips_max_norm2(self._h, C.byref(out), self._stream()))
        return float(out.value)

    @property
    def phi(self) -> float:
        return float(self._L.mips_get_phi(self._h))

    @phi.setter
    def phi(self, v: float) -> None:
        check(self._L.mips_set_phi(self._h, float(v)))

    def reconstruct_n(self, i0: int = 0, n: Optional[int] = None, as_torch: bool = False):
        """Stored rows [i0, i0+n) as float32 (bf16 banks return the rounded values)."""
        n = self.ntotal - i0 if n is None else n
        if as_torch:
            out = torch.empty((n, self.d), dtype=torch.float32, device=self.device)
            check(self._L.mips_reconstruct(self._h, i0, n, _ptr(out), 1, self._stream()))
            return out
        out = np.empty((n, self.d), dtype=np.float32)
        check(self._L.mips_reconstruct(self._h, i0, n, out.ctypes.data_as(C.c_void_p), 0, self._stream()))
        return out

    @staticmethod
    def _algo_code(algo: str) -> int:
        return _ALGOS[algo]

    def gather_rows(self, ids: torch.Tensor) -> torch.Tensor:
        """Stored rows by GLOBAL id, float32 [..., d] on the device: what the reference
        re-encodes per step (mips.py:465-470) when the memory encoder is frozen. ids outside this
        shard (other ranks' rows, -1 padding) give zero rows — sum the ranks' outputs for a sharded bank."""
        if not ids.is_cuda or ids.dtype != torch.int64:
            raise ValueError("gather_rows needs a CUDA int64 id tensor")
        flat = ids.contiguous().view(-1)
        out = torch.empty((flat.shape[0], self.d), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._L.mips_gather_rows(self._h, _ptr(flat), flat.shape[0], self.id_offset, _ptr(out), self._stream()))
        return out.view(*ids.shape, self.d)

    # ------------------------------------------------------------------ search side
    def _check_k(self, k: int, multipass: bool = False) -> int:
        """One search pass keeps up to MAX_K (64) results per query; the public search calls accept up to
        MAX_K_MULTIPASS by running bounded passes (faiss accepts any k: mips.py:383-386)."""
        k = int(k)
        top = MAX_K_MULTIPASS if multipass else MAX_K
        if k < 1 or k > top:
            raise ValueError(f"k must be in [1, {top}], got {k}")
        return k

    def search(self, xq, k: int, return_torch: bool = False):
        """faiss Index.search(xq, k) -> (D float32 [nq,k], I int64 [nq,k]); IP scores descending,
        L2 squared distances ascending, -1 padding. numpy in -> numpy out through the host entry
        point (H2D + kernels + D2H in one C call); CUDA tensor in (or return_torch=True) keeps
        everything on the device (removes the round trip of retriever_generator.py:143)."""
        k = self._check_k(k, multipass=True)
        if isinstance(xq, torch.Tensor) and (xq.is_cuda or return_torch):
            r = self.search_ex(xq, k)
            return r["scores"], r["ids"]
        if isinstance(xq, torch.Tensor):
            xq = xq.detach().cpu().numpy()
        xq = np.asarray(xq)
        if xq.ndim != 2:
            raise ValueError("Shape of query must be 2D")
        if xq.shape[1] != self.d:
            raise ValueError(f"Query vectors must have dimension {self.d}, got {xq.shape[1]}")
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        nq = xq.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        out_mode = OUT_IP if self.metric_type == METRIC_INNER_PRODUCT else OUT_L2
        with torch.cuda.device(self.device):
            check(self._L.mips_search_host(self._h, xq.ctypes.data_as(C.c_void_p), nq, k, 0, None, out_mode,
                                           D.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p),
                                           self._stream()))
        if self.id_offset:
            I = np.where(I >= 0, I + self.id_offset, I)
        return D, I

    def search_host(self, xq: np.ndarray, k: int, ignore_ids: Optional[np.ndarray] = None,
                    normalize_queries: bool = False, out_mode: Optional[int] = None,
                    D: Optional[np.ndarray] = None, I: Optional[np.ndarray] = None):
        """The reference-facing end-to-end call (host buffers in and out, one C-ABI call):
        Mips.search semantics incl. the per-query ignored id (mips.py:382-400). ids are local
        (id_offset is not applied). D / I may be preallocated (e.g. pinned) arrays. k > 64 runs bounded
        passes inside the C call."""
        k = self._check_k(k, multipass=True)
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        if xq.ndim != 2 or xq.shape[1] != self.d:
            raise ValueError(f"Shape of query must be [nq, {self.d}], got {xq.shape}")
        nq = xq.shape[0]
        D = np.empty((nq, k), dtype=np.float32) if D is None else D
        I = np.empty((nq, k), dtype=np.int64) if I is None else I
        ign = None
        if ignore_ids is not None:
            ign_arr = np.ascontiguousarray(ignore_ids, dtype=np.int64)
            if ign_arr.shape != (nq,):
                raise ValueError("ignore_ids must have one id per query")
            ign = ign_arr.ctypes.data_as(C.c_void_p)
        if out_mode is None:
            out_mode = OUT_IP if self.metric_type == METRIC_INNER_PRODUCT else OUT_L2
        with torch.cuda.device(self.device):
            check(self._L.mips_search_host(self._h, xq.ctypes.data_as(C.c_void_p), nq, k,
                                           int(normalize_queries), ign, int(out_mode),
                                           D.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p),
                                           self._stream()))
        return D, I

    def search_local(self, xq: torch.Tensor, k: int, ignore_ids: Optional[torch.Tensor] = None,
                     normalize_queries: bool = False, algo: str = "auto", after=None):
        """K1 + local merge on this shard: returns device tensors
        (key [nq,k] descending ranking key, ids [nq,k] GLOBAL int64, xnorm2 [nq,k], qnorm2 [nq]).
        after = (key [nq], id [nq]): one pass of a multi-pass search — only rows strictly after that
        (key, id) in the order (key descending, id ascending) are eligible (mips_search_local_after)."""
        k = self._check_k(k)
        if not isinstance(xq, torch.Tensor):
            xq = torch.as_tensor(np.ascontiguousarray(xq, dtype=np.float32))
        if xq.dim() != 2:
            raise ValueError("Shape of query must be 2D")
        if xq.shape[1] != self.d:
            raise ValueError(f"Query vectors must have dimension {self.d}, got {xq.shape[1]}")
        xq = xq.detach().to(device=self.device, dtype=torch.float32).contiguous()
        nq = xq.shape[0]
        key = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        xn2 = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        qn2 = torch.empty((nq,), dtype=torch.float32, device=self.device)
        ign = None
        if ignore_ids is not None:
            ign = torch.as_tensor(ignore_ids).to(device=self.device, dtype=torch.int64).contiguous()
            if ign.shape != (nq,):
                raise ValueError("ignore_ids must have one id per query")
        a_key = a_id = None
        if after is not None:
            a_key = after[0].to(device=self.device, dtype=torch.float32).contiguous()
            a_id = after[1].to(device=self.device, dtype=torch.int64).contiguous()
        with torch.cuda.device(self.device):
            check(self._L.mips_search_local_after(self._h, _ptr(xq), nq, k, int(normalize_queries), _ptr(ign),
                                                  self.id_offset, _ALGOS[algo], _ptr(a_key), _ptr(a_id), _ptr(key),
                                                  _ptr(ids), _ptr(xn2), _ptr(qn2), None, self._stream()))
        return key, ids, xn2, qn2

    def search_local_multipass(self, xq: torch.Tensor, k: int, ignore_ids: Optional[torch.Tensor] = None,
                               normalize_queries: bool = False, algo: str = "auto"):
        """search_local for any k <= MAX_K_MULTIPASS: passes of MAX_K results, each bounded by the last result of
        the pass before (exact: the concatenation is the shard's top-k in (key desc, id asc) order)."""
        k = self._check_k(k, multipass=True)
        if self.dtype == "fp32" and algo == "auto" and k > MAX_K:
            algo = "simt"        # the bound compares keys for equality: every pass on the same (exact fp32) kernel
        keys, idss, xn2s, qn2, after = [], [], [], None, None
        for done in range(0, k, MAX_K):
            kp = min(MAX_K, k - done)
            key, ids, xn2, qn2 = self.search_local(xq, kp, ignore_ids=ignore_ids, normalize_queries=normalize_queries,
                                                   algo=algo, after=after)
            keys.append(key), idss.append(ids), xn2s.append(xn2)
            last_id = ids[:, -1]
            after = (torch.where(last_id < 0, torch.full_like(key[:, -1], float("-inf")), key[:, -1]),
                     torch.where(last_id < 0, torch.full_like(last_id, torch.iinfo(torch.int64).max), last_id))
        return torch.cat(keys, 1), torch.cat(idss, 1), torch.cat(xn2s, 1), qn2

    def search_local_packed(self, xq: torch.Tensor, k: int, ignore_ids: Optional[torch.Tensor] = None,
                            normalize_queries: bool = False, algo: str = "auto"):
        """K1 + local merge with the result as 16-byte records {f32 key, f32 |x|^2, i64 global id}:
        returns (packed uint8 [nq, k, 16], qnorm2 [nq]) — the buffer one all-gather moves."""
        k = self._check_k(k)
        if not isinstance(xq, torch.Tensor):
            xq = torch.as_tensor(np.ascontiguousarray(xq, dtype=np.float32))
        if xq.dim() != 2:
            raise ValueError("Shape of query must be 2D")
        if xq.shape[1] != self.d:
            raise ValueError(f"Query vectors must have dimension {self.d}, got {xq.shape[1]}")
        xq = xq.detach().to(device=self.device, dtype=torch.float32).contiguous()
        nq = xq.shape[0]
        packed = torch.empty((nq, k, 16), dtype=torch.uint8, device=self.device)
        qn2 = torch.empty((nq,), dtype=torch.float32, device=self.device)
        ign = None
        if ignore_ids is not None:
            ign = torch.as_tensor(ignore_ids).to(device=self.device, dtype=torch.int64).contiguous()
            if ign.shape != (nq,):
                raise ValueError("ignore_ids must have one id per query")
        with torch.cuda.device(self.device):
            check(self._L.mips_search_local_packed(self._h, _ptr(xq), nq, k, int(normalize_queries), _ptr(ign),
                                                   self.id_offset, _ALGOS[algo], _ptr(packed), _ptr(qn2),
                                                   self._stream()))
        return packed, qn2

    def merge(self, key: torch.Tensor, ids: torch.Tensor, xn2: Optional[torch.Tensor], qn2: Optional[torch.Tensor],
              k: int, want: Iterable[str] = ("scores", "ids"), out_mode: Optional[int] = None,
              ignore_ids: Optional[torch.Tensor] = None, mem_len: Optional[int] = None,
              beta: float = 1.0, beta_bias: float = 0.0) -> dict:
        """K2 over candidate lists [n_parts, nq, k_in] (one part per shard after the all-gather, or a
        single local list): final (D, I) + the doc-score outputs of retriever_generator.py:158-193."""
        return merge_candidates(key, ids, xn2, qn2, k, self.metric_type, want=want, out_mode=out_mode,
                                phi=self.phi, ignore_ids=ignore_ids, mem_len=mem_len, beta=beta,
                                beta_bias=beta_bias)

    def merge_packed(self, packed: torch.Tensor, qn2: Optional[torch.Tensor], k: int,
                     want: Iterable[str] = ("scores", "ids"), out_mode: Optional[int] = None,
                     mem_len: Optional[int] = None, beta: float = 1.0, beta_bias: float = 0.0) -> dict:
        return merge_candidates(None, None, None, qn2, k, self.metric_type, want=want, out_mode=out_mode,
                                phi=self.phi, mem_len=mem_len, beta=beta, beta_bias=beta_bias, packed=packed)

    def search_ex(self, xq, k: int, ignore_ids=None, want: Iterable[str] = ("scores", "ids"),
                  L: Optional[int] = None, normalize_queries: bool = False, out_mode: Optional[int] = None,
                  beta: float = 1.0, beta_bias: float = 0.0, algo: str = "auto") -> dict:
        """Device-resident search with optional fused outputs. want ⊆ {"scores","ids","cosine",
        "doc_prob","memory_bias"}; L = memory_seq_len for memory_bias. k > 64: bounded passes, outputs from
        the same arithmetic spelled with tensor ops (finalize_lists; a rare path)."""
        if int(k) > MAX_K:
            key, ids, xn2, qn2 = self.search_local_multipass(xq, k, ignore_ids=ignore_ids,
                                                             normalize_queries=normalize_queries, algo=algo)
            return finalize_lists(key, ids, xn2, qn2, self.metric_type, out_mode, self.phi, want, L, beta, beta_bias)
        key, ids, xn2, qn2 = self.search_local(xq, k, ignore_ids=ignore_ids,
                                               normalize_queries=normalize_queries, algo=algo)
        return self.merge(key.unsqueeze(0), ids.unsqueeze(0), xn2.unsqueeze(0), qn2, k, want=want,
                          out_mode=out_mode, mem_len=L, beta=beta, beta_bias=beta_bias)

    def capture(self, nq: int, k: int, with_ignore: bool = False, want: Iterable[str] = ("scores", "ids"),
                L: Optional[int] = None, normalize_queries: bool = False, out_mode: Optional[int] = None,
                beta: float = 1.0, beta_bias: float = 0.0, algo: str = "auto", host_io: bool = False) -> "GraphedSearch":
        """Capture query prep -> K1 -> merges (one C-ABI call, mips_search_sharded with one rank) as ONE CUDA
        graph over static buffers; a replay is a single launch. The graph holds the addresses of the bank shard
        and of the handle's scratch: capture the LARGEST (nq, k) shape first (a later, larger search regrows the
        scratch and would leave an earlier graph pointing at freed memory), and capture again after `add` grew
        the shard or after a refresh."""
        k = self._check_k(k)
        return GraphedSearch(self, int(nq), k, with_ignore, want, L,
                             lambda xq, ign, out: sharded_step(self, None, 1, 0, False, xq, ign, k, out, L,
                                                               normalize_queries, out_mode, beta, beta_bias, algo),
                             host_io=host_io)

    # ------------------------------------------------------------------ profiling hooks (bench)
    def set_profiling(self, on: bool) -> None:
        check(self._L.mips_set_profiling(self._h, int(on)))

    def k1_ms_total(self):
        return float(self._L.mips_k1_ms_total(self._h)), int(self._L.mips_prof_count(self._h))

    @property
    def last_algo(self) -> str:
        return self._L.mips_last_algo(self._h).decode()

    def fallback_queries(self, reset: bool = True) -> int:
        """Exact tensor-core search ("tcx", fp32 banks): how many queries failed the exactness
        certificate since the last reset and were recomputed by the exact SIMT kernel."""
        return int(self._L.mips_fallback_queries(self._h, int(reset)))


def finalize_lists(key, ids, xn2, qn2, metric_type: int, out_mode, phi: float, want, L, beta: float,
                   beta_bias: float) -> dict:
    """The output transform of the merge kernel (k2_merge.cuh, FINAL mode) for result lists longer than one kernel
    pass holds (k > 64, multi-pass searches): ranking key -> inner product -> D per out_mode, cosine doc scores
    (retriever_generator.py:159-172), per-doc softmax, memory_bias (:188-192). [nq, k] elementwise tensor ops."""
    want = set(want)
    if out_mode is None:
        out_mode = OUT_IP if metric_type == METRIC_INNER_PRODUCT else OUT_L2
    valid = ids >= 0
    ip = key + 0.5 * xn2 if metric_type == METRIC_L2 else key
    q2 = qn2[:, None]
    if out_mode == OUT_IP:
        D = torch.where(valid, ip, torch.full_like(ip, float("-inf")))
    else:
        dist_ = torch.clamp(q2 + xn2 - 2.0 * ip, min=0.0) if out_mode == OUT_L2 else q2 + float(phi) - 2.0 * ip
        D = torch.where(valid, dist_, torch.full_like(ip, float("inf")))
    out = {"scores": D, "ids": ids}
    if want & {"cosine", "doc_prob", "memory_bias"}:
        den = torch.sqrt(q2) * torch.sqrt(xn2)
        cos = torch.where(valid & (den > 0), ip / den.clamp_min(1e-38), torch.zeros_like(ip))
        out["cosine"] = cos
        if "doc_prob" in want:
            z = torch.where(valid, beta * cos + beta_bias, torch.full_like(cos, float("-inf")))
            out["doc_prob"] = torch.nan_to_num(torch.softmax(z, dim=1), nan=0.0)
        if "memory_bias" in want:
            if not L or L < 1:
                raise ValueError("memory_bias needs L (memory_seq_len) >= 1")
            out["memory_bias"] = torch.where(valid, cos, torch.zeros_like(cos)).repeat_interleave(int(L), dim=1)
    return out


def alloc_outputs(dev, nq: int, k: int, want, L) -> dict:
    """Output tensors of one search step for the requested `want` set."""
    want = set(want)
    unknown = want - {"scores", "ids", "cosine", "doc_prob", "memory_bias"}
    if unknown:
        raise ValueError(f"unknown outputs requested: {sorted(unknown)}")
    out = {"scores": torch.empty((nq, k), dtype=torch.float32, device=dev),
           "ids": torch.empty((nq, k), dtype=torch.int64, device=dev)}
    if want & {"cosine", "memory_bias", "doc_prob"}:
        out["cosine"] = torch.empty((nq, k), dtype=torch.float32, device=dev)
    if "doc_prob" in want:
        out["doc_prob"] = torch.empty((nq, k), dtype=torch.float32, device=dev)
    if "memory_bias" in want:
        if not L or L < 1:
            raise ValueError("memory_bias needs L (memory_seq_len) >= 1")
        out["memory_bias"] = torch.empty((nq, k * int(L)), dtype=torch.float32, device=dev)
    return out


def sharded_step(local: "B200FlatIndex", comm_ptr, world: int, rank: int, dp: bool, xq: torch.Tensor,
                 ign: Optional[torch.Tensor], k: int, out: dict, L, normalize_queries, out_mode, beta, beta_bias,
                 algo: str) -> dict:
    """ONE C-ABI call for a whole search step on the current stream: mips_search_sharded (queries replicated)
    or mips_search_sharded_dp (every rank its own queries). world == 1 needs no communicator."""
    lib = _lib.lib()
    if out_mode is None:
        out_mode = OUT_IP if local.metric_type == METRIC_INNER_PRODUCT else OUT_L2
    with torch.cuda.device(local.device):
        tail = (int(out_mode), _ptr(out["scores"]), _ptr(out["ids"]), _ptr(out.get("cosine")), _ptr(out.get("doc_prob")),
                float(beta), float(beta_bias), _ptr(out.get("memory_bias")), int(L or 0), local._stream())
        if dp:
            check(lib.mips_search_sharded_dp(local._h, comm_ptr, world, rank, _ptr(xq), xq.shape[0], int(k),
                                             int(normalize_queries), _ptr(ign), local.id_offset, _ALGOS[algo], *tail))
        else:
            check(lib.mips_search_sharded(local._h, comm_ptr, world, _ptr(xq), xq.shape[0], int(k),
                                          int(normalize_queries), _ptr(ign), local.id_offset, _ALGOS[algo], *tail))
    return out


class GraphedSearch:
    """One CUDA graph holding a whole search step (B200FlatIndex.capture, ShardedFlatIndex.capture). `xq` (and
    `ignore_ids`) are static input tensors: replay(xq) copies into them (from the device or from pinned host
    memory) and launches the graph; results land in `.out` (static, overwritten by the next replay)."""

    def __init__(self, local: "B200FlatIndex", nq: int, k: int, with_ignore: bool, want, L, call, host_io: bool = False):
        dev = local.device
        self.nq, self.k = nq, k
        # random placeholder queries: all-zero queries tie every row of the bank, the worst case of the exact
        # fp32 search (every query fails its certificate and takes the slow fallback during the warm-up)
        self.xq = torch.randn((nq, local.d), dtype=torch.float32, device=dev)
        self.ignore_ids = torch.full((nq,), -1, dtype=torch.int64, device=dev) if with_ignore else None
        self.out = alloc_outputs(dev, nq, k, want, L)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):             # warm-up: scratch reaches its final size, NCCL connects its peers
            for _ in range(2):
                call(self.xq, self.ignore_ids, self.out)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # host_io: the graph also holds the H2D copy of the queries from a pinned host buffer (`xq_host`: write the
        # step's queries there) and the D2H copies of every output into pinned host tensors (`out_host`) — the
        # whole host-to-host step is ONE launch (replay_host)
        self.xq_host = self.out_host = None
        if host_io:
            self.xq_host = torch.empty((nq, local.d), dtype=torch.float32).pin_memory()
            self.xq_host.copy_(self.xq)
            self.out_host = {name: torch.empty(t.shape, dtype=t.dtype).pin_memory() for name, t in self.out.items()}
        self._call, self._host_io = call, host_io
        self.replicas = []
        self.graph = self._capture_one()

    def _capture_one(self):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if self._host_io:
                self.xq.copy_(self.xq_host, non_blocking=True)
            self._call(self.xq, self.ignore_ids, self.out)
            if self._host_io:
                for name, t in self.out.items():
                    self.out_host[name].copy_(t, non_blocking=True)
        return g

    def add_replicas(self, n: int) -> list:
        """n more graphs of the same step over the same static buffers. With profiling on (set_profiling) every
        capture records its own pair of timing events around K1, so a loop over [graph] + replicas leaves one K1
        duration per step (bench.py's roofline measurement inside the timed region)."""
        for _ in range(n):
            self.replicas.append(self._capture_one())
        return [self.graph] + self.replicas

    def close(self) -> None:
        """Destroy the graph (required before the NCCL communicator it captured can be destroyed)."""
        if self.graph is not None:
            torch.cuda.synchronize(self.xq.device)
            for g in [self.graph] + self.replicas:
                g.reset()
            self.graph, self.replicas = None, []

    def replay_host(self, xq=None) -> dict:
        """host_io graphs: queries from the pinned `xq_host` (optionally filled from `xq` first), results in the
        pinned `out_host` tensors; returns after the device is done (one launch + one synchronize)."""
        if self.graph is None or self.xq_host is None:
            raise RuntimeError("capture(..., host_io=True) first")
        if xq is not None:
            self.xq_host.copy_(torch.as_tensor(xq))
        self.graph.replay()
        torch.cuda.current_stream(self.xq.device).synchronize()
        return self.out_host

    def replay(self, xq: Optional[torch.Tensor] = None, ignore_ids: Optional[torch.Tensor] = None) -> dict:
        if self.graph is None:
            raise RuntimeError("this captured search was closed (its index was refreshed or closed)")
        if xq is not None:
            self.xq.copy_(xq, non_blocking=True)
        if ignore_ids is not None:
            self.ignore_ids.copy_(ignore_ids, non_blocking=True)
        self.graph.replay()
        return self.out


def merge_candidates(key: Optional[torch.Tensor], ids: Optional[torch.Tensor], xn2: Optional[torch.Tensor],
                     qn2: Optional[torch.Tensor], k: int, metric_type: int,
                     want: Iterable[str] = ("scores", "ids"), out_mode: Optional[int] = None,
                     phi: float = 0.0, ignore_ids: Optional[torch.Tensor] = None,
                     mem_len: Optional[int] = None, beta: float = 1.0, beta_bias: float = 0.0,
                     packed: Optional[torch.Tensor] = None) -> dict:
    """K2. Candidates either as three arrays [n_parts, nq, k_in] (key, ids, xn2) or as one packed
    uint8 tensor [n_parts, nq, k_in, 16] of {f32 key, f32 |x|^2, i64 id} records."""
    L = _lib.lib()
    want = set(want)
    unknown = want - {"scores", "ids", "cosine", "doc_prob", "memory_bias"}
    if unknown:
        raise ValueError(f"unknown outputs requested: {sorted(unknown)}")
    if packed is not None:
        if packed.dim() != 4 or packed.shape[-1] != 16 or packed.dtype != torch.uint8:
            raise ValueError("packed candidates must be uint8 [n_parts, nq, k_in, 16]")
        src = packed = packed.contiguous()
        n_parts, nq, k_in = packed.shape[:3]
    else:
        if key.dim() != 3 or key.shape != ids.shape:
            raise ValueError("candidates must be [n_parts, nq, k_in]")
        src = key
        n_parts, nq, k_in = key.shape
        key = key.contiguous()
        ids = ids.contiguous()
        xn2 = None if xn2 is None else xn2.contiguous()
    if not src.is_cuda:
        raise RuntimeError("merge runs on the GPU; candidates must be CUDA tensors")
    dev = src.device
    if out_mode is None:
        out_mode = OUT_IP if metric_type == METRIC_INNER_PRODUCT else OUT_L2
    D = torch.empty((nq, k), dtype=torch.float32, device=dev)
    I = torch.empty((nq, k), dtype=torch.int64, device=dev)
    cosine = torch.empty((nq, k), dtype=torch.float32, device=dev) if want & {"cosine", "memory_bias", "doc_prob"} else None
    doc_prob = torch.empty((nq, k), dtype=torch.float32, device=dev) if "doc_prob" in want else None
    mbias = None
    if "memory_bias" in want:
        if not mem_len or mem_len < 1:
            raise ValueError("memory_bias needs L (memory_seq_len) >= 1")
        mbias = torch.empty((nq, k * int(mem_len)), dtype=torch.float32, device=dev)
    ign = None if ignore_ids is None else torch.as_tensor(ignore_ids).to(device=dev, dtype=torch.int64).contiguous()
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if packed is not None:
            check(L.mips_merge_packed(_ptr(packed), n_parts, nq, k_in, int(k), int(metric_type), int(out_mode),
                                      float(phi), _ptr(qn2), _ptr(ign), _ptr(D), _ptr(I), _ptr(cosine),
                                      _ptr(doc_prob), float(beta), float(beta_bias), _ptr(mbias),
                                      int(mem_len or 0), st))
        else:
            check(L.mips_merge(_ptr(key), _ptr(ids), _ptr(xn2), n_parts, nq, k_in, int(k), int(metric_type),
                               int(out_mode), float(phi), _ptr(qn2), _ptr(ign), _ptr(D), _ptr(I), _ptr(cosine),
                               _ptr(doc_prob), float(beta), float(beta_bias), _ptr(mbias), int(mem_len or 0), st))
    out = {"scores": D, "ids": I}
    if cosine is not None:
        out["cosine"] = cosine
    if doc_prob is not None:
        out["doc_prob"] = doc_prob
    if mbias is not None:
        out["memory_bias"] = mbias
    return out


# ---------------------------------------------------------------------------------------------
# faiss-module-shaped helpers (what the reference imports from `faiss` for this path)
def IndexFlatIP(d: int, **kw) -> B200FlatIndex:
    return B200FlatIndex(d, METRIC_INNER_PRODUCT, **kw)


def IndexFlatL2(d: int, **kw) -> B200FlatIndex:
    return B200FlatIndex(d, METRIC_L2, **kw)


def IndexFlat(d: int, metric: int = METRIC_L2, **kw) -> B200FlatIndex:
    return B200FlatIndex(d, metric, **kw)


def index_factory(d: int, description: str = "Flat", metric: int = METRIC_L2, **kw) -> B200FlatIndex:
    """faiss.index_factory for the exact path only (mips_string_factory "Flat",
    model_config.py:50). Approximate factories (IVF*, HNSW*, SQ8 ...; sotasum/config.yaml:94)
    are out of scope and rejected loudly."""
    if description.strip() != "Flat":
        raise ValueError(f"only the exact 'Flat' factory is supported, got {description!r}")
    return B200FlatIndex(d, metric, **kw)


def normalize_L2(x) -> None:
    """faiss.normalize_L2(x): in-place row normalisation (mips.py:521-525), run on the GPU."""
    L = _lib.lib()
    if isinstance(x, torch.Tensor) and x.is_cuda:
        if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 2:
            raise ValueError("normalize_L2 needs a contiguous float32 [n, d] tensor")
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        check(L.mips_normalize_l2(_ptr(x), x.shape[0], x.shape[1], 1, x.device.index, st))
        return
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or not x.flags.c_contiguous or x.ndim != 2:
        raise ValueError("normalize_L2 needs a C-contiguous float32 [n, d] array")
    dev = _device_index(None)
    check(L.mips_normalize_l2(x.ctypes.data_as(C.c_void_p), x.shape[0], x.shape[1], 0, dev, None))


def retriever_metrics(ids: torch.Tensor, row_aid: torch.Tensor, query_aid: torch.Tensor, counts: torch.Tensor,
                      return_pred: bool = False) -> dict:
    """retriever_metrics (pretrain.py:69-85) + the hit matrix of mips.py:456-463, on the device:
    ids int64 [B, k] as returned by the search, row_aid int64 [N] (the `aid` of every memory row),
    query_aid int64 [B], counts float32 [B] (`aid_counts`). Returns python floats like the reference
    (one 12-byte read back), plus the per-query terms and optionally the hit matrix as CUDA tensors."""
    if not (ids.is_cuda and row_aid.is_cuda and query_aid.is_cuda and counts.is_cuda):
        raise ValueError("retriever_metrics runs on the GPU: pass CUDA tensors (no CPU compute path)")
    if ids.dim() != 2 or ids.dtype != torch.int64:
        raise ValueError("ids must be int64 [B, k]")
    B, k = ids.shape
    dev = ids.device
    ids, row_aid = ids.contiguous(), row_aid.to(torch.int64).contiguous()
    query_aid, counts = query_aid.to(torch.int64).contiguous(), counts.to(torch.float32).contiguous()
    per_q = torch.empty((B, 3), dtype=torch.float32, device=dev)
    out3 = torch.zeros(3, dtype=torch.float32, device=dev)
    pred = torch.empty((B, k), dtype=torch.float32, device=dev) if return_pred else None
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(_lib.lib().mips_retriever_metrics(_ptr(ids), B, k, _ptr(row_aid), row_aid.shape[0], _ptr(query_aid),
                                                _ptr(counts), _ptr(per_q), _ptr(out3), _ptr(pred), st))
    r, rr, ap = out3.cpu().tolist()
    out = {"recall": r, "reciprocal_rank": rr, "average_precision": ap, "per_query": per_q}
    if return_pred:
        out["pred"] = pred
    return out


class MemoryTokenStore:
    """The memory's documents tokenised ONCE and kept in HBM ([N, L] int32 + token counts), gathered by the
    search's ids on the device. Replaces the per-step host work of reference
    sotasum/mips.py:428,473-501: Arrow text lookup of the retrieved rows, `memory_tokenizer(flat_texts,
    padding="max_length", truncation=True)`, and the derived masks."""

    def __init__(self, input_ids, lengths=None, attention_mask=None, pad_id: int = 1, bos_id: int = 0,
                 eos_id: int = 2, device=None):
        dev = torch.device("cuda", _device_index(device))
        ids = torch.as_tensor(np.asarray(input_ids) if not isinstance(input_ids, torch.Tensor) else input_ids)
        if ids.dim() != 2:
            raise ValueError("input_ids must be [N, L]")
        if lengths is None:
            if attention_mask is None:
                raise ValueError("pass lengths [N] or attention_mask [N, L]")
            am = torch.as_tensor(np.asarray(attention_mask) if not isinstance(attention_mask, torch.Tensor) else attention_mask)
            lengths = am.to(torch.int64).sum(1)         # right padding (padding="max_length"): a prefix of ones
        self.input_ids = ids.to(dev, torch.int32).contiguous()
        self.lengths = torch.as_tensor(lengths).to(dev, torch.int32).contiguous()
        self.pad_id, self.bos_id, self.eos_id = int(pad_id), int(bos_id), int(eos_id)
        self.device = dev

    @property
    def n_rows(self) -> int:
        return int(self.input_ids.shape[0])

    @property
    def seq_len(self) -> int:
        return int(self.input_ids.shape[1])

    def gather(self, ids: torch.Tensor) -> dict:
        """ids int64 [B, k] (CUDA) -> memory_input_ids / attention_mask / memory_attention_mask /
        global_attention_mask, int64 [B*k, L] like the tensors of mips.py:473-501."""
        if not ids.is_cuda or ids.dtype != torch.int64:
            raise ValueError("gather needs a CUDA int64 id tensor")
        flat = ids.contiguous().view(-1)
        n, L = flat.shape[0], self.seq_len
        outs = [torch.empty((n, L), dtype=torch.int64, device=self.device) for _ in range(4)]
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            check(_lib.lib().mips_gather_tokens(_ptr(self.input_ids), _ptr(self.lengths), self.n_rows, L, _ptr(flat), n,
                                                self.pad_id, self.bos_id, self.eos_id, *[_ptr(o) for o in outs], st))
        return {"memory_input_ids": outs[0], "attention_mask": outs[1], "memory_attention_mask": outs[2],
                "global_attention_mask": outs[3]}
