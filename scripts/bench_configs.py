"""Secondary measurements for the other BASELINE.json configs (not the headline bench line):
C2 exactness config (1M x 768 fp32, 256 queries, k=8), C4 retriever-generator step shape
(16 queries, k=5, fused cosine / doc_prob / memory_bias), C5 refresh (K0 ingest GB/s) + k=32 search.
Prints one JSON object per config. Single GPU."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m

dev = torch.device("cuda", 0)
PEAK_HBM = 6528.4


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def build(n, d, dtype, seed=1):
    idx = m.B200FlatIndex(d, 0, dtype=dtype, capacity=n)
    gen = torch.Generator(device=dev).manual_seed(seed)
    t_add = 0.0
    for s in range(0, n, 500_000):
        blk = torch.randn((min(500_000, n - s), d), generator=gen, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx.add(blk)
        e1.record()
        torch.cuda.synchronize()
        t_add += e0.elapsed_time(e1)
    return idx, t_add


out = []
# ---- C2
n, d, nq, k = 1_000_000, 768, 256, 8
idx, t_add = build(n, d, "fp32")
xq = torch.randn((nq, d), device=dev)
ms = timed(lambda: idx.search_ex(xq, k), iters=5, warm=2)
idx.fallback_queries(reset=True)
idx.search_ex(xq, k)
out.append({"config": "C2 1Mx768 fp32 nq=256 k=8", "kernel": idx.last_algo, "ms": ms, "qps": nq / ms * 1e3,
            "fallback_queries": idx.fallback_queries(), "hbm_roofline_ms_fp32_rows": n * d * 4 / PEAK_HBM / 1e6,
            "hbm_roofline_ms_bf16_shadow": n * d * 2 / PEAK_HBM / 1e6, "add_gbs": (n * d * 10) / t_add / 1e6})
ms = timed(lambda: idx.search_ex(xq, k, algo="simt"), iters=5, warm=2)
out.append({"config": "C2 1Mx768 fp32 nq=256 k=8 (plain fp32-FMA kernel)", "kernel": idx.last_algo, "ms": ms,
            "qps": nq / ms * 1e3, "tflops_fp32": 2 * nq * n * d / ms / 1e9})
del idx
# ---- C2 on a bf16 bank (tensor path) for comparison
idx, t_add = build(n, d, "bf16")
ms = timed(lambda: idx.search_ex(xq, k), iters=20)
out.append({"config": "C2-shape 1Mx768 bf16 nq=256 k=8", "kernel": idx.last_algo, "ms": ms, "qps": nq / ms * 1e3,
            "hbm_gbs": n * d * 2 / ms / 1e6, "hbm_frac": n * d * 2 / ms / 1e6 / PEAK_HBM})
del idx
# ---- C4 / C5 on a 10M bf16 bank
n = 10_000_000
idx, t_add = build(n, d, "bf16")
out.append({"config": "K0 ingest 10Mx768 fp32->bf16 (device rows)", "ms": t_add,
            "gbs_read_plus_write": n * d * 6 / t_add / 1e6, "hbm_frac": n * d * 6 / t_add / 1e6 / PEAK_HBM})
for nq_, k_, want, L in ((16, 5, ("scores", "ids", "cosine", "doc_prob", "memory_bias"), 512), (1, 8, ("scores", "ids"), None),
                         (128, 8, ("scores", "ids"), None), (256, 8, ("scores", "ids"), None), (1024, 32, ("scores", "ids"), None)):
    xq = torch.randn((nq_, d), device=dev)
    ms = timed(lambda: idx.search_ex(xq, k_, want=want, L=L), iters=20)
    out.append({"config": f"10Mx768 bf16 nq={nq_} k={k_} want={'+'.join(want)}", "kernel": idx.last_algo, "ms": ms,
                "qps": nq_ / ms * 1e3, "hbm_gbs": n * d * 2 / ms / 1e6, "hbm_frac": n * d * 2 / ms / 1e6 / PEAK_HBM,
                "tflops": 2 * nq_ * n * d / ms / 1e9})
for o in out:
    print(json.dumps(o))
