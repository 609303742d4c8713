"""BASELINE config 5 — memory refresh: rebuild a 2M-doc index from freshly "encoded" embeddings that
are already on the GPU (the encoder itself is out of scope: SURVEY §8 A7/A8), then search throughput
at k=32, row-sharded over the ranks of one box.

    python scripts/bench_c5.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/bench_c5.py

Replaces the reference's refresh protocol (lightning_model.py:148-180: 3 barriers, a disk round
trip of the whole bank, three 1000-row-batch host passes and a rank-0 index.add) by K0 into the
rank's HBM shard + two tiny collectives (row counts, phi)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import retrieval_augmented_mds_b200 as m
from retrieval_augmented_mds_b200.sharded import ShardedFlatIndex, balanced_range

N, D, NQ, K = 2_000_000, 768, 1024, 32
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
import faulthandler
faulthandler.dump_traceback_later(600, exit=True)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows = balanced_range(N, rank, world)
gen = torch.Generator(device=dev).manual_seed(99 + rank)
emb = torch.randn((len(rows), D), generator=gen, device=dev)          # "encoder output" of this rank's documents
xq = torch.randn((NQ, D), generator=torch.Generator(device="cpu").manual_seed(5)).to(dev)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


idx = m.B200FlatIndex(D, m.METRIC_INNER_PRODUCT, dtype="bf16", device=dev, capacity=len(rows))
sh = ShardedFlatIndex(idx) if world > 1 else None


def rebuild():
    """The refresh: drop the rows, keep the HBM allocation, ingest this rank's new embeddings, agree
    on the id space and on phi (Mips.build_index + _build_mips_index2 of the reference)."""
    idx.reset()
    if sh is not None:
        sh.add_local(emb)
        sh.sync_phi()
    else:
        idx.add(emb)
        idx.phi = idx.max_norm2()


rebuild()                                                            # warm-up (allocator, NCCL)
sync()
t0 = time.perf_counter()
rebuild()
sync()
rebuild_ms = 1e3 * (time.perf_counter() - t0)
search = (lambda: sh.search(xq, K)) if sh is not None else (lambda: idx.search_ex(xq, K))
steps = 50


def timed(fn):
    for _ in range(5):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


eager_ms = timed(search)
idx.set_profiling(True)
for _ in range(20):
    search()
torch.cuda.synchronize()
k1_ms, k1_n = idx.k1_ms_total()
idx.set_profiling(False)
g = sh.capture(NQ, K) if sh is not None else idx.capture(NQ, K)      # the whole step as ONE CUDA graph
g.xq.copy_(xq)
graph_ms = timed(lambda: g.graph.replay())
t = torch.tensor([graph_ms, eager_ms, rebuild_ms, k1_ms / max(k1_n, 1)], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms, eg, rb, k1 = (float(v) for v in t)
    bound_ms = 2.0 * NQ * N * D / world / 1686.4e9
    print(json.dumps({"config": f"C5 rebuild {N}x{D} index (device embeddings) + search nq={NQ} k={K}", "n_gpus": world,
                      "rebuild_ms": rb, "rebuild_rows_per_s": N / rb * 1e3, "search_ms": ms, "search_ms_eager": eg,
                      "k1_ms": k1, "search_qps": NQ / ms * 1e3, "kernel": idx.last_algo,
                      "search_tflops": 2.0 * NQ * N * D / ms / 1e9, "tensor_bound_ms": bound_ms,
                      "frac_of_bound": bound_ms / ms, "step": "one CUDA graph replay (max over ranks)"}))
if sh is not None:
    sh.close()
if world > 1:
    dist.destroy_process_group()
