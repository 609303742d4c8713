"""Exact fp32 search, small batches on a large bank (the C4 shape on an fp32 memory): 4M x 768 fp32."""
import json, sys
import torch
sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m
dev = torch.device("cuda:0")
n, d = 4_000_000, 768
idx = m.B200FlatIndex(d, 0, dtype="fp32", capacity=n)
gen = torch.Generator(device=dev).manual_seed(1)
for s in range(0, n, 500_000):
    idx.add(torch.randn((500_000, d), generator=gen, device=dev))
for nq, k in ((16, 5), (128, 8), (256, 8)):
    xq = torch.randn((nq, d), generator=gen, device=dev)
    for _ in range(3):
        idx.search_ex(xq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        idx.search_ex(xq, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"config": f"{n}x{d} fp32 exact nq={nq} k={k}", "kernel": idx.last_algo, "ms": ms,
                      "fallback_queries": idx.fallback_queries(), "shadow_pass_hbm_ms": n * d * 2 / 6528.4e6}))
