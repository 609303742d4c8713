"""BASELINE config 4 under data-parallel training — every rank searches with its OWN batch of 16 query CLS
vectors (as each DDP rank of the reference calls self.mips(queries=...), retriever_generator.py:143-153), k=5,
fused cosine / doc_prob / memory_bias outputs (L=512), against the 10M x 768 bf16 bank row-sharded over the ranks:
all-gather queries -> one local search of G*16 queries -> all-to-all of 16-byte records -> per-rank merge.

    python scripts/bench_c4_dp.py                                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29514 scripts/bench_c4_dp.py
"""
import faulthandler
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import retrieval_augmented_mds_b200 as m
from retrieval_augmented_mds_b200.sharded import ShardedFlatIndex, balanced_range

faulthandler.dump_traceback_later(600, exit=True)
N, D, B, K, L = 10_000_000, 768, 16, 5, 512
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows = balanced_range(N, rank, world)
idx = m.B200FlatIndex(D, m.METRIC_INNER_PRODUCT, dtype="bf16", device=dev, capacity=len(rows), id_offset=rows.start)
gen = torch.Generator(device=dev).manual_seed(99 + rank)
for s in range(0, len(rows), 500_000):
    idx.add(torch.randn((min(500_000, len(rows) - s), D), generator=gen, device=dev))
xq = torch.randn((B, D), generator=gen, device=dev)                  # this rank's own queries
want = ("scores", "ids", "cosine", "doc_prob", "memory_bias")
sh = None
if world > 1:
    sh = ShardedFlatIndex(idx)
    cnt = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    allc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, cnt)
    sh.counts = [int(c) for c in allc.cpu().tolist()]
    step = lambda: sh.search_dp(xq, K, want=want, L=L)
    g = sh.capture(B, K, dp=True, want=want, L=L)
else:
    step = lambda: idx.search_ex(xq, K, want=want, L=L)
    g = idx.capture(B, K, want=want, L=L)
g.xq.copy_(xq)


def timed(fn, steps=50):
    for _ in range(5):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


eager = timed(step)
graph = timed(lambda: g.graph.replay())
a, b = step(), g.replay(xq)
torch.cuda.synchronize()
same = bool(torch.equal(a["ids"], b["ids"]))
t = torch.tensor([graph, eager], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    hbm_bound = len(rows) * D * 2 / 6528.4e6
    print(json.dumps({"config": f"C4 data-parallel step: {world} rank(s) x {B} own queries, k={K}, 10M x {D} bf16 bank sharded, "
                                f"fused cosine / doc_prob / memory_bias (L={L})", "n_gpus": world, "step_ms_graph": float(t[0]),
                      "step_ms_eager": float(t[1]), "queries_per_s": world * B / float(t[0]) * 1e3, "kernel": idx.last_algo,
                      "hbm_bound_ms_per_shard": hbm_bound, "frac_of_hbm_bound": hbm_bound / float(t[0]),
                      "graph_equals_eager": same}))
if sh is not None:
    sh.close()
if world > 1:
    dist.destroy_process_group()
