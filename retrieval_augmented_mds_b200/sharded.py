"""Row-sharded bank across the GPUs of one box: one process per GPU, each rank owns a contiguous
row range (the partition of Mips.encode_text2, sotasum/mips.py:226-230), searches it locally and
the per-rank top-k lists are combined by ONE all-gather followed by the on-device merge (K2).

The reference never shards: rank 0 builds and every rank loads a full CPU replica
(lightning_model.py:168-180). Top-k over a union of row sets is the merge of per-set top-k, so
the sharded result is exact and independent of the number of shards.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import METRIC_IP as METRIC_INNER_PRODUCT, OUT_IP, OUT_L2


def shard_range(n_rows: int, rank: int, num_rank: int) -> range:
    """Rows owned by `rank` under the reference's partition rule (mips.py:226-230)."""
    chunk = (n_rows // num_rank) + 1
    stop = (rank + 1) * chunk if rank + 1 < num_rank else n_rows
    start = min(rank * chunk, n_rows)
    return range(start, max(min(stop, n_rows), start))


def balanced_range(n_rows: int, rank: int, num_rank: int) -> range:
    """ceil(N/G) partition used for synthetic benches."""
    chunk = (n_rows + num_rank - 1) // num_rank
    return range(min(rank * chunk, n_rows), min((rank + 1) * chunk, n_rows))


def weighted_ranges(n_rows: int, weights, max_skew: float = 0.10) -> list:
    """Contiguous row ranges proportional to per-rank `weights` (measured search throughput): boards of
    one box differ by several percent under the power cap, and a strong-scaling step waits for the
    slowest shard. Each share is clamped to (1 +- max_skew) x the equal share; ranges tile [0, n_rows)."""
    w = [max(float(x), 0.0) for x in weights]
    g = len(w)
    if g == 0 or sum(w) <= 0.0:
        raise ValueError("weights must be positive")
    mean = sum(w) / g
    w = [min(max(x, mean * (1.0 - max_skew)), mean * (1.0 + max_skew)) for x in w]
    tot = sum(w)
    bounds = [0]
    acc = 0.0
    for r in range(g):
        acc += w[r]
        bounds.append(n_rows if r == g - 1 else min(n_rows, int(round(n_rows * acc / tot))))
    return [range(bounds[r], max(bounds[r + 1], bounds[r])) for r in range(g)]


def exchange_offsets(n_local: int, group=None, device=None):
    """All ranks learn every shard's row count; returns (id_offset_of_this_rank, counts list)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    counts = [int(v) for v in out.cpu().tolist()]
    return sum(counts[:rank]), counts


def allreduce_max(value: float, group=None, device=None) -> float:
    """Global phi / max_norm^2: MAX of one float per rank (replaces the rank-0 pass of
    mips.py:298-304, 316-324)."""
    t = torch.tensor([value], dtype=torch.float32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_candidates(key: torch.Tensor, ids: torch.Tensor, xn2: torch.Tensor, group=None):
    """One all-gather of the packed per-rank lists. key/xn2 fp32 [nq,k], ids int64 [nq,k] ->
    ([G,nq,k] key, [G,nq,k] ids, [G,nq,k] xn2). The three arrays travel as ONE int32 buffer
    [nq, k, 4] (key bits, xn2 bits, id lo, id hi) = 16 bytes per candidate."""
    world = dist.get_world_size(group)
    nq, k = key.shape
    packed = torch.empty((nq, k, 4), dtype=torch.int32, device=key.device)
    packed[..., 0] = key.contiguous().view(torch.int32)
    packed[..., 1] = xn2.contiguous().view(torch.int32)
    packed[..., 2:4] = ids.contiguous().view(torch.int32).view(nq, k, 2)
    out = torch.empty((world * nq, k, 4), dtype=torch.int32, device=key.device)
    dist.all_gather_into_tensor(out, packed, group=group)   # concatenation along dim 0, rank major
    out = out.view(world, nq, k, 4)
    g_key = out[..., 0].contiguous().view(torch.float32)
    g_xn2 = out[..., 1].contiguous().view(torch.float32)
    g_ids = out[..., 2:4].contiguous().view(torch.int64).view(world, nq, k)
    return g_key, g_ids, g_xn2


def gather_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """THE exchange step of the search path: one all-gather of each rank's packed top-k list
    (uint8 [nq, k, 16] -> [G, nq, k, 16]); 16*nq*k bytes per rank (128 KiB at nq=1024, k=8)."""
    world = dist.get_world_size(group)
    nq, k, rec = packed.shape
    out = torch.empty((world * nq, k, rec), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return out.view(world, nq, k, rec)


class PeerExchange:
    """Exchange buffers of the peer-memory variant of the cross-GPU step (include/mips_b200.h, "peer-memory
    exchange"): every rank cudaMallocs one buffer, the ranks swap CUDA IPC handles once and map each
    other's buffers; afterwards a search needs no collective at all — the local merge kernel stores this
    rank's records into every rank's buffer over NVLink and the final merge kernel waits on arrival flags.
    Two slot sets alternate between consecutive searches. Creation and close() are collective."""

    FLAG_BYTES = 128          # one slot's flags: up to 32 ranks x uint32

    def __init__(self, device: torch.device, group, nq_cap: int, k_cap: int):
        self.L = _lib.lib()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 32:
            raise ValueError("peer-memory exchange supports up to 32 ranks")
        self.device = device
        self.nq_cap, self.k_cap = int(nq_cap), int(k_cap)
        self.slot_bytes = (self.world * self.nq_cap * self.k_cap * 16 + 255) // 256 * 256
        self.flags_off = 2 * self.slot_bytes
        total = self.flags_off + 2 * self.FLAG_BYTES
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(self.L.mips_xchg_alloc(device.index, total, C.byref(ptr), handle))
        self.own = int(ptr.value)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.peers = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                self.peers.append(self.own)
                continue
            p = C.c_void_p()
            _lib.check(self.L.mips_xchg_open(device.index, C.create_string_buffer(hb, 64), C.byref(p)))
            self.peers.append(int(p.value))
        self.seq = 0
        self._ptr_cache = {}
        dist.barrier(group)

    def fits(self, nq: int, k: int) -> bool:
        return self.world * nq * k * 16 <= self.slot_bytes

    def next_search(self, nq: int, k: int):
        """-> (seq, device array of this rank's regions on every rank, device array of its flags there,
        address of this rank's own slot, address of its flags) for the next search of shape (nq, k)."""
        self.seq += 1
        slot = self.seq & 1
        key = (nq, k, slot)
        if key not in self._ptr_cache:
            mine = slot * self.slot_bytes + self.rank * nq * k * 16
            bufs = torch.tensor([p + mine for p in self.peers], dtype=torch.int64, device=self.device)
            flags = torch.tensor([p + self.flags_off + slot * self.FLAG_BYTES + 4 * self.rank for p in self.peers],
                                 dtype=torch.int64, device=self.device)
            self._ptr_cache[key] = (bufs, flags)
        bufs, flags = self._ptr_cache[key]
        return (self.seq, bufs, flags, self.own + slot * self.slot_bytes,
                self.own + self.flags_off + slot * self.FLAG_BYTES)

    def close(self) -> None:
        if self.own is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                       # nobody is still writing into a buffer about to go
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.L.mips_xchg_close(self.device.index, C.c_void_p(p))
        dist.barrier(self.group)
        self.L.mips_xchg_free(self.device.index, C.c_void_p(self.own))
        self.own = None


class ShardedFlatIndex:
    """B200FlatIndex per rank + NCCL all-gather + K2 merge. Queries are replicated on all ranks
    (every rank passes the same xq) and every rank ends with the full result."""

    def __init__(self, local_index, group=None, exchange: Optional[str] = None):
        self.local = local_index
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.counts = [0] * self.world
        # "nccl": one all-gather of the packed lists; "p2p": peer-memory exchange fused into the merge
        # kernels (CUDA IPC over NVLink, no collective launch per search); MIPS_B200_EXCHANGE overrides
        self.exchange = exchange or os.environ.get("MIPS_B200_EXCHANGE", "nccl")
        if self.exchange not in ("nccl", "p2p"):
            raise ValueError(f"exchange must be 'nccl' or 'p2p', got {self.exchange!r}")
        self._xchg: Optional[PeerExchange] = None

    @property
    def d(self) -> int:
        return self.local.d

    @property
    def metric_type(self) -> int:
        return self.local.metric_type

    @property
    def ntotal(self) -> int:
        return int(sum(self.counts))

    def add_local(self, x, normalize: bool = False) -> None:
        """Append this rank's rows, then agree on the id space: global id = offset[rank] + row.
        Collective (every rank must call it)."""
        self.local.add(x, normalize=normalize)
        off, self.counts = exchange_offsets(self.local.ntotal, self.group, self.local.device)
        self.local.id_offset = off

    def sync_phi(self) -> float:
        """phi = max over ALL shards of |x|^2 (get_phi, mips.py:55-56). Collective."""
        phi = allreduce_max(self.local.max_norm2(), self.group, self.local.device)
        self.local.phi = phi
        return phi

    def search(self, xq, k: int, ignore_ids=None, want: Iterable[str] = ("scores", "ids"),
               L: Optional[int] = None, normalize_queries: bool = False, out_mode: Optional[int] = None,
               beta: float = 1.0, beta_bias: float = 0.0, algo: str = "auto") -> dict:
        if self.exchange == "p2p" and self.world > 1 and self.local.dtype == "bf16":
            r = self._search_p2p(xq, k, ignore_ids, set(want), L, normalize_queries, out_mode, beta, beta_bias, algo)
            if r is not None:
                return r
        packed, qn2 = self.local.search_local_packed(xq, k, ignore_ids=ignore_ids,
                                                     normalize_queries=normalize_queries, algo=algo)
        gathered = gather_packed(packed, self.group) if self.world > 1 else packed.unsqueeze(0)
        return self.local.merge_packed(gathered, qn2, k, want=want, out_mode=out_mode, mem_len=L,
                                       beta=beta, beta_bias=beta_bias)

    # ------------------------------------------------------------------ peer-memory exchange
    def _search_p2p(self, xq, k, ignore_ids, want, L, normalize_queries, out_mode, beta, beta_bias, algo):
        loc = self.local
        lib = _lib.lib()
        if not isinstance(xq, torch.Tensor):
            xq = torch.as_tensor(xq)
        if xq.dim() != 2 or xq.shape[1] != loc.d:
            raise ValueError(f"Query vectors must be [nq, {loc.d}]")
        xq = xq.detach().to(device=loc.device, dtype=torch.float32).contiguous()
        nq, k = xq.shape[0], int(k)
        if nq == 0 or nq > 148 * 128:
            return None                                   # chunked searches keep the NCCL path
        if self._xchg is None or not self._xchg.fits(nq, k):
            if self._xchg is not None:
                self._xchg.close()
            self._xchg = PeerExchange(loc.device, self.group, max(nq, 1024), max(k, 8))
        seq, bufs, flags, my_buf, my_flags = self._xchg.next_search(nq, k)
        dev = loc.device
        qn2 = torch.empty((nq,), dtype=torch.float32, device=dev)
        ign = None if ignore_ids is None else torch.as_tensor(ignore_ids).to(device=dev, dtype=torch.int64).contiguous()
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        cosine = torch.empty((nq, k), dtype=torch.float32, device=dev) if want & {"cosine", "memory_bias", "doc_prob"} else None
        doc_prob = torch.empty((nq, k), dtype=torch.float32, device=dev) if "doc_prob" in want else None
        mbias = None
        if "memory_bias" in want:
            if not L or L < 1:
                raise ValueError("memory_bias needs L (memory_seq_len) >= 1")
            mbias = torch.empty((nq, k * int(L)), dtype=torch.float32, device=dev)
        if out_mode is None:
            out_mode = OUT_IP if loc.metric_type == METRIC_INNER_PRODUCT else OUT_L2
        vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.mips_search_local_xchg(loc._h, vp(xq), nq, k, int(normalize_queries), vp(ign), loc.id_offset,
                                                  loc._algo_code(algo), vp(bufs), vp(flags), self.world, seq, vp(qn2), st))
            _lib.check(lib.mips_merge_xchg(C.c_void_p(my_buf), C.c_void_p(my_flags), self.world, seq, nq, k, k,
                                           loc.metric_type, int(out_mode), float(loc.phi), vp(qn2), None, vp(D), vp(I),
                                           vp(cosine), vp(doc_prob), float(beta), float(beta_bias), vp(mbias),
                                           int(L or 0), st))
        out = {"scores": D, "ids": I}
        if cosine is not None:
            out["cosine"] = cosine
        if doc_prob is not None:
            out["doc_prob"] = doc_prob
        if mbias is not None:
            out["memory_bias"] = mbias
        return out

    def close(self) -> None:
        """Collective: release the peer-memory exchange buffers (no-op for the NCCL exchange)."""
        if self._xchg is not None:
            self._xchg.close()
            self._xchg = None
