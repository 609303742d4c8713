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
import weakref
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib
from . import index as _index
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


def alltoall_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """The exchange step of the data-parallel TRAINING search through torch.distributed (the C-ABI NCCL path
    does the same with one ncclSend/ncclRecv group): `packed` uint8 [G * B, k, 16] holds this shard's lists
    for every rank's B queries, rank major; rank r receives [G, B, k, 16] = every shard's lists of ITS queries."""
    world = dist.get_world_size(group)
    gb, k, rec = packed.shape
    if gb % world:
        raise ValueError("first dimension must be world * B")
    out = torch.empty_like(packed)
    dist.all_to_all_single(out, packed.contiguous(), group=group)
    return out.view(world, gb // world, k, rec)


class NcclComm:
    """An NCCL communicator owned by libmips_b200 (include/mips_b200.h, "cross-GPU step through NCCL"): rank 0
    draws the unique id through the C ABI, the existing torch.distributed group (any backend) carries its 128
    bytes to the other ranks, and every rank calls ncclCommInitRank through the C ABI. Afterwards the search
    step never touches torch.distributed: the collective is enqueued by the C side on the caller's stream."""

    def __init__(self, device: torch.device, group=None):
        self.L = _lib.lib()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.ptr = C.c_void_p()
        if self.L.mips_nccl_version() < 0:
            raise _lib.MipsError("NCCL is not available to libmips_b200 (libnccl.so.2 not loadable)")
        box = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            _lib.check(self.L.mips_nccl_unique_id(buf))
            box[0] = bytes(buf.raw)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
        with torch.cuda.device(device):
            _lib.check(self.L.mips_nccl_comm_init(C.byref(self.ptr), self.world, self.rank,
                                                  C.create_string_buffer(box[0], 128), device.index))

    def close(self) -> None:
        if self.ptr is not None and self.ptr.value:
            torch.cuda.synchronize(self.device)
            self.L.mips_nccl_comm_destroy(self.ptr)
            self.ptr = None


class PeerExchange:
    """Exchange buffers of the peer-memory variant of the cross-GPU step (include/mips_b200.h, "peer-memory
    exchange"): every rank cudaMallocs one buffer, the ranks swap CUDA IPC handles once and map each
    other's buffers; afterwards a search needs no collective at all — the local merge kernel stores this
    rank's records into every rank's buffer over NVLink and the final merge kernel waits on arrival flags.
    Two slot sets alternate between consecutive searches. Creation and close() are collective."""

    FLAG_BYTES = 256          # one slot's flags: 32 arrival flags (uint32) + the time-out flag (word 32)

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

    def timed_out(self) -> int:
        """Sequence number of the last search whose wait for a peer expired (0 = none). Synchronises."""
        worst = 0
        for slot in (0, 1):
            v = C.c_uint32(0)
            _lib.check(self.L.mips_xchg_timeout_seq(self.device.index,
                                                    C.c_void_p(self.own + self.flags_off + slot * self.FLAG_BYTES),
                                                    C.byref(v)))
            worst = max(worst, int(v.value))
        return worst

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
    """B200FlatIndex per rank + one collective + K2 merge.

    `search`: queries are replicated on all ranks (every rank passes the same xq) and every rank ends with
    the full result. `search_dp`: the data-parallel TRAINING step — each rank passes its OWN batch, as every
    DDP rank of the reference calls `self.mips(queries=...)` with its own queries
    (sotasum/retriever_generator.py:143-153, driven per rank by lightning_model.py:188-216), and gets the
    global top-k of its own queries.

    exchange = "native" (default on NCCL groups): the whole step is ONE C-ABI call; the collective is a raw
    ncclAllGather (or ncclSend/ncclRecv group) enqueued by libmips_b200 on the caller's stream over its own
    communicator — no torch.distributed dispatch in the step, capturable in a CUDA graph (`capture`).
    "torch": torch.distributed collectives between the C-ABI calls (any backend; the first round's path).
    "p2p": peer-memory exchange fused into the merge kernels (CUDA IPC over NVLink)."""

    def __init__(self, local_index, group=None, exchange: Optional[str] = None):
        self.local = local_index
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.counts = [0] * self.world
        default = "native" if dist.get_backend(group) == "nccl" else "torch"
        self.exchange = exchange or os.environ.get("MIPS_B200_EXCHANGE", default)
        if self.exchange == "nccl":           # round-1 name of the torch.distributed path
            self.exchange = "torch"
        if self.exchange not in ("native", "torch", "p2p"):
            raise ValueError(f"exchange must be 'native', 'torch' or 'p2p', got {self.exchange!r}")
        self._xchg: Optional[PeerExchange] = None
        self._comm: Optional[NcclComm] = None
        self._graphs = []            # weak references to the CUDA graphs captured over the communicator

    @property
    def d(self) -> int:
        return self.local.d

    @property
    def metric_type(self) -> int:
        return self.local.metric_type

    @property
    def ntotal(self) -> int:
        return int(sum(self.counts))

    def adopt(self, other: "ShardedFlatIndex") -> None:
        """Take over the communicator / exchange buffers of `other` (the index this one replaces after a
        memory refresh) instead of creating new ones; `other` must not be used afterwards (graphs captured
        over it point at the old shard and are released)."""
        other._release_graphs()
        self.exchange = other.exchange
        self._comm, other._comm = other._comm, None
        self._xchg, other._xchg = other._xchg, None

    def _release_graphs(self) -> None:
        for ref in self._graphs:
            g = ref()
            if g is not None:
                g.close()
        self._graphs = []

    def comm(self) -> NcclComm:
        """The library-owned NCCL communicator (created collectively on first use)."""
        if self._comm is None:
            self._comm = NcclComm(self.local.device, self.group)
        return self._comm

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

    # ------------------------------------------------------------------ outputs
    def _alloc_outputs(self, nq: int, k: int, want, L):
        return _index.alloc_outputs(self.local.device, nq, k, want, L)

    def _prep(self, xq, ignore_ids):
        loc = self.local
        if not isinstance(xq, torch.Tensor):
            xq = torch.as_tensor(xq)
        if xq.dim() != 2:
            raise ValueError("Shape of query must be 2D")
        if xq.shape[1] != loc.d:
            raise ValueError(f"Query vectors must have dimension {loc.d}, got {xq.shape[1]}")
        xq = xq.detach().to(device=loc.device, dtype=torch.float32).contiguous()
        ign = None
        if ignore_ids is not None:
            ign = torch.as_tensor(ignore_ids).to(device=loc.device, dtype=torch.int64).contiguous()
            if ign.shape != (xq.shape[0],):
                raise ValueError("ignore_ids must have one id per query")
        return xq, ign

    def _native_call(self, dp: bool, xq, ign, k, out, L, normalize_queries, out_mode, beta, beta_bias, algo):
        """ONE C-ABI call for the whole sharded step (mips_search_sharded / mips_search_sharded_dp)."""
        comm = self.comm().ptr if self.world > 1 else None
        return _index.sharded_step(self.local, comm, self.world, self.rank, dp, xq, ign, k, out, L,
                                   normalize_queries, out_mode, beta, beta_bias, algo)

    # ------------------------------------------------------------------ k > 64 (rare: bounded passes per shard)
    def _search_big_k(self, dp: bool, xq, k, ignore_ids, want, L, normalize_queries, out_mode, beta, beta_bias, algo):
        """k > 64: every shard runs its bounded passes (B200FlatIndex.search_local_multipass), the [nq, k] lists
        travel through torch.distributed and are merged by (key desc, id asc) with two stable sorts."""
        xq, ign = self._prep(xq, ignore_ids)
        B = xq.shape[0]
        if dp and self.world > 1:
            q_all = torch.empty((self.world * B, xq.shape[1]), dtype=torch.float32, device=xq.device)
            dist.all_gather_into_tensor(q_all, xq, group=self.group)
            if ign is not None:
                ign_all = torch.empty((self.world * B,), dtype=torch.int64, device=xq.device)
                dist.all_gather_into_tensor(ign_all, ign, group=self.group)
                ign = ign_all
            xq = q_all
        key, ids, xn2, qn2 = self.local.search_local_multipass(xq, k, ignore_ids=ign, normalize_queries=normalize_queries,
                                                               algo=algo)
        if self.world > 1:
            nq = xq.shape[0]
            packed = torch.stack([key.double(), xn2.double(), ids.double()], 0)      # ids < 2^53: exact in fp64
            allp = torch.empty((self.world * 3, nq, k), dtype=torch.float64, device=xq.device)
            dist.all_gather_into_tensor(allp, packed, group=self.group)    # concatenation along dim 0, rank major
            allp = allp.view(self.world, 3, nq, k)
            ck = allp[:, 0].permute(1, 0, 2).reshape(nq, -1).float()
            cx = allp[:, 1].permute(1, 0, 2).reshape(nq, -1).float()
            ci = allp[:, 2].permute(1, 0, 2).reshape(nq, -1).long()
            o1 = torch.sort(torch.where(ci < 0, torch.full_like(ci, torch.iinfo(torch.int64).max), ci), dim=1, stable=True)[1]
            ck, cx, ci = ck.gather(1, o1), cx.gather(1, o1), ci.gather(1, o1)
            o2 = torch.sort(ck, dim=1, descending=True, stable=True)[1][:, :k]
            key, xn2, ids = ck.gather(1, o2), cx.gather(1, o2), ci.gather(1, o2)
            if dp:
                mine = slice(self.rank * B, (self.rank + 1) * B)
                key, xn2, ids, qn2 = key[mine], xn2[mine], ids[mine], qn2[mine]
        return _index.finalize_lists(key, ids, xn2, qn2, self.local.metric_type, out_mode, self.local.phi, want, L, beta,
                                     beta_bias)

    # ------------------------------------------------------------------ replicated queries
    def search(self, xq, k: int, ignore_ids=None, want: Iterable[str] = ("scores", "ids"),
               L: Optional[int] = None, normalize_queries: bool = False, out_mode: Optional[int] = None,
               beta: float = 1.0, beta_bias: float = 0.0, algo: str = "auto") -> dict:
        k = self.local._check_k(k, multipass=True)
        if k > _lib.MAX_K:
            return self._search_big_k(False, xq, k, ignore_ids, want, L, normalize_queries, out_mode, beta, beta_bias, algo)
        if self.exchange == "p2p" and self.world > 1 and self.local.dtype == "bf16":
            r = self._search_p2p(xq, k, ignore_ids, set(want), L, normalize_queries, out_mode, beta, beta_bias, algo)
            if r is not None:
                return r
        if self.exchange in ("native", "p2p"):
            xq, ign = self._prep(xq, ignore_ids)
            out = self._alloc_outputs(xq.shape[0], k, want, L)
            return self._native_call(False, xq, ign, k, out, L, normalize_queries, out_mode, beta, beta_bias, algo)
        packed, qn2 = self.local.search_local_packed(xq, k, ignore_ids=ignore_ids,
                                                     normalize_queries=normalize_queries, algo=algo)
        gathered = gather_packed(packed, self.group) if self.world > 1 else packed.unsqueeze(0)
        return self.local.merge_packed(gathered, qn2, k, want=want, out_mode=out_mode, mem_len=L,
                                       beta=beta, beta_bias=beta_bias)

    # ------------------------------------------------------------------ data-parallel training step
    def search_dp(self, xq_local, k: int, ignore_ids=None, want: Iterable[str] = ("scores", "ids"),
                  L: Optional[int] = None, normalize_queries: bool = False, out_mode: Optional[int] = None,
                  beta: float = 1.0, beta_bias: float = 0.0, algo: str = "auto") -> dict:
        """Each rank passes its OWN [B, d] queries (same B on every rank; ignore_ids given on all ranks or on
        none) and receives the global top-k of its own queries: all-gather queries -> one local search of
        G*B queries -> all-to-all of 16-byte records -> per-rank merge with the fused doc outputs
        (SURVEY §8e "DP-training variant"). Collective."""
        k = self.local._check_k(k, multipass=True)
        if k > _lib.MAX_K:
            return self._search_big_k(True, xq_local, k, ignore_ids, want, L, normalize_queries, out_mode, beta, beta_bias,
                                      algo)
        xq, ign = self._prep(xq_local, ignore_ids)
        B = xq.shape[0]
        if self.exchange in ("native", "p2p") or self.world == 1:
            out = self._alloc_outputs(B, k, want, L)
            return self._native_call(True, xq, ign, k, out, L, normalize_queries, out_mode, beta, beta_bias, algo)
        q_all = torch.empty((self.world * B, xq.shape[1]), dtype=torch.float32, device=xq.device)
        dist.all_gather_into_tensor(q_all, xq, group=self.group)
        ign_all = None
        if ign is not None:
            ign_all = torch.empty((self.world * B,), dtype=torch.int64, device=xq.device)
            dist.all_gather_into_tensor(ign_all, ign, group=self.group)
        packed, qn2 = self.local.search_local_packed(q_all, k, ignore_ids=ign_all,
                                                     normalize_queries=normalize_queries, algo=algo)
        mine = alltoall_packed(packed, self.group)
        return self.local.merge_packed(mine, qn2[self.rank * B:(self.rank + 1) * B].contiguous(), k, want=want,
                                       out_mode=out_mode, mem_len=L, beta=beta, beta_bias=beta_bias)

    # ------------------------------------------------------------------ CUDA graph of the whole step
    def capture(self, nq: int, k: int, dp: bool = False, with_ignore: bool = False,
                want: Iterable[str] = ("scores", "ids"), L: Optional[int] = None, normalize_queries: bool = False,
                out_mode: Optional[int] = None, beta: float = 1.0, beta_bias: float = 0.0,
                algo: str = "auto", host_io: bool = False) -> "GraphedSearch":
        """Capture query prep -> K1 -> local merge -> NCCL collective -> final merge as ONE CUDA graph over
        static buffers (collective: every rank captures). Replays cost one launch; results land in
        `.out` (static tensors, overwritten by the next replay). Capture the largest (nq, k) shape first: the
        graph holds the addresses of the handle's scratch, which a later, larger search regrows (see
        B200FlatIndex.capture); a refresh (`adopt`) and `close()` release the graphs."""
        if self.exchange not in ("native", "p2p") and self.world > 1:
            raise ValueError("capture needs the native exchange (the C-ABI NCCL step)")
        k = self.local._check_k(k)
        g = _index.GraphedSearch(
            self.local, int(nq), k, with_ignore, want, L,
            lambda xq, ign, out: self._native_call(dp, xq, ign, k, out, L, normalize_queries, out_mode, beta,
                                                   beta_bias, algo),
            host_io=host_io)
        self._graphs.append(weakref.ref(g))
        return g

    # ------------------------------------------------------------------ peer-memory exchange
    def check_exchange(self) -> None:
        """Raise if a peer-memory search timed out waiting for a peer (its queries came back with ids -1);
        the index switches to the NCCL exchange for the following searches."""
        if self._xchg is not None and self._xchg.timed_out():
            self.exchange = "native" if dist.get_backend(self.group) == "nccl" else "torch"
            raise _lib.MipsError("peer-memory exchange: a peer's list did not arrive in time; "
                                 "falling back to the NCCL exchange")

    def _search_p2p(self, xq, k, ignore_ids, want, L, normalize_queries, out_mode, beta, beta_bias, algo):
        loc = self.local
        lib = _lib.lib()
        xq, ign = self._prep(xq, ignore_ids)
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
        out = self._alloc_outputs(nq, k, want, L)
        if out_mode is None:
            out_mode = OUT_IP if loc.metric_type == METRIC_INNER_PRODUCT else OUT_L2
        vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.mips_search_local_xchg(loc._h, vp(xq), nq, k, int(normalize_queries), vp(ign), loc.id_offset,
                                                  loc._algo_code(algo), vp(bufs), vp(flags), self.world, seq, vp(qn2), st))
            _lib.check(lib.mips_merge_xchg(C.c_void_p(my_buf), C.c_void_p(my_flags), self.world, seq, nq, k, k,
                                           loc.metric_type, int(out_mode), float(loc.phi), vp(qn2), None,
                                           vp(out["scores"]), vp(out["ids"]), vp(out.get("cosine")),
                                           vp(out.get("doc_prob")), float(beta), float(beta_bias),
                                           vp(out.get("memory_bias")), int(L or 0), st))
        return out

    def close(self) -> None:
        """Collective: release the captured graphs, the peer-memory exchange buffers and the library-owned
        communicator. The graphs go FIRST: NCCL keeps a communicator alive (ncclCommDestroy blocks) while a
        CUDA graph that captured one of its collectives exists."""
        self._release_graphs()
        if self._xchg is not None:
            self._xchg.close()
            self._xchg = None
        if self._comm is not None:
            self._comm.close()
            self._comm = None
