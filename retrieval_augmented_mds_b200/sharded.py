"""Row-sharded bank across the GPUs of one box: one process per GPU, each rank owns a contiguous
row range (the partition of Mips.encode_text2, sotasum/mips.py:226-230), searches it locally and
the per-rank top-k lists are combined by ONE all-gather followed by the on-device merge (K2).

The reference never shards: rank 0 builds and every rank loads a full CPU replica
(lightning_model.py:168-180). Top-k over a union of row sets is the merge of per-set top-k, so
the sharded result is exact and independent of the number of shards.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


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


class ShardedFlatIndex:
    """B200FlatIndex per rank + NCCL all-gather + K2 merge. Queries are replicated on all ranks
    (every rank passes the same xq) and every rank ends with the full result."""

    def __init__(self, local_index, group=None):
        self.local = local_index
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.counts = [0] * self.world

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
        packed, qn2 = self.local.search_local_packed(xq, k, ignore_ids=ignore_ids,
                                                     normalize_queries=normalize_queries, algo=algo)
        gathered = gather_packed(packed, self.group) if self.world > 1 else packed.unsqueeze(0)
        return self.local.merge_packed(gathered, qn2, k, want=want, out_mode=out_mode, mem_len=L,
                                       beta=beta, beta_bias=beta_bias)
