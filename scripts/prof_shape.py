"""One search shape, for timing (CUDA events, eager and CUDA-graph) or for an ncu capture of its kernels.
    python scripts/prof_shape.py --rows 250000 --nq 1024 --k 32 [--dtype bf16|fp32] [--iters 50] [--ncu]
Prints one JSON line. --ncu: 2 warm-up searches + 1 search only (capture with `ncu -k regex:search_ --launch-skip`)."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, required=True)
ap.add_argument("--nq", type=int, required=True)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--algo", default="auto")
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
idx = m.B200FlatIndex(a.d, 0, dtype=a.dtype, capacity=a.rows)
gen = torch.Generator(device=dev).manual_seed(1)
for s in range(0, a.rows, 500_000):
    idx.add(torch.randn((min(500_000, a.rows - s), a.d), generator=gen, device=dev))
xq = torch.randn((a.nq, a.d), generator=gen, device=dev)
if a.ncu:
    for _ in range(3):
        idx.search_ex(xq, a.k, algo=a.algo)
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(5):
    idx.search_ex(xq, a.k, algo=a.algo)
torch.cuda.synchronize()
idx.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    idx.search_ex(xq, a.k, algo=a.algo)
e1.record()
torch.cuda.synchronize()
k1_ms, k1_n = idx.k1_ms_total()
idx.set_profiling(False)
eager = e0.elapsed_time(e1) / a.iters
g = idx.capture(a.nq, min(a.k, 64), algo=a.algo)
g.xq.copy_(xq)
for _ in range(5):
    g.graph.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(a.iters):
    g.graph.replay()
e1.record()
torch.cuda.synchronize()
graph = e0.elapsed_time(e1) / a.iters
eb = 2 if a.dtype == "bf16" else 4
print(json.dumps({"tag": a.tag, "rows": a.rows, "nq": a.nq, "k": a.k, "d": a.d, "dtype": a.dtype, "algo": idx.last_algo,
                  "eager_ms": eager, "graph_ms": graph, "k1_ms": k1_ms / max(k1_n, 1),
                  "hbm_bound_ms": a.rows * a.d * eb / 6528.4e6, "tensor_bound_ms": 2.0 * a.nq * a.rows * a.d / 1686.4e9,
                  "fallbacks": idx.fallback_queries() if a.dtype == "fp32" else None}))
