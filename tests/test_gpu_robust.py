"""Robustness of the search path: randomised shapes against the float64 oracle (every K1 variant,
both metrics, ignore ids), and CUDA-graph capture of a whole search (no hidden allocation or host
synchronisation after warm-up: SURVEY §8b 'ownership')."""
import numpy as np
import pytest
import torch

import retrieval_augmented_mds_b200 as pkg
from oracle import mips_oracle as o

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_match_oracle(cuda_device, seed):
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice([64, 72, 128, 200, 384, 512, 576, 768, 1000, 1024]))
    n = int(rng.integers(1, 9000))
    nq = int(rng.integers(1, 600))
    k = int(rng.choice([1, 2, 5, 8, 13, 16, 17, 32, 40, 64]))
    metric = int(rng.integers(0, 2))
    xb = rng.standard_normal((n, d), dtype=np.float32) * rng.uniform(0.3, 3.0, (n, 1)).astype(np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    if n > 50:                                   # a cluster of exact duplicates and near-duplicates
        xb[n // 3: n // 3 + 9] = xb[0]
        xb[n // 2: n // 2 + 5] = xb[1] * (1 + 1e-6)
        xq[: min(nq, 4)] = xb[0] * 1.5
    ign = rng.integers(0, n, nq).astype(np.int64) if seed % 2 else None
    for dtype in ("bf16", "fp32"):
        idx = pkg.B200FlatIndex(d, metric, dtype=dtype)
        idx.add(xb)
        stored = xb if dtype == "fp32" else o.bf16_round(xb)
        qq = xq if dtype == "fp32" else o.bf16_round(xq)
        D_ref, I_ref = o.exact_topk_f64(stored, qq, k, metric, ignore=ign)
        algos = ["auto", "simt"] + (["tc2"] if dtype == "bf16" and (k <= 32 or d <= 768) else []) + \
                (["tc"] if dtype == "bf16" and d <= 768 else [])
        for algo in algos:
            r = idx.search_ex(torch.from_numpy(xq), k, algo=algo,
                              ignore_ids=None if ign is None else torch.from_numpy(ign))
            torch.cuda.synchronize()
            o.check_topk(stored, qq, r["scores"].cpu().numpy(), r["ids"].cpu().numpy(), metric,
                         rtol=1e-5 if dtype == "fp32" else 1e-4, D_ref=D_ref, I_ref=I_ref, ignore=ign,
                         what=f"seed {seed} {dtype} {algo} n={n} d={d} nq={nq} k={k} m={metric}")
        idx.close()


@pytest.mark.parametrize("dtype,nq", [("bf16", 300), ("bf16", 16), ("fp32", 200)])
def test_search_is_cuda_graph_capturable(cuda_device, dtype, nq):
    """One search = K0 + K1 (+ K3) + K2 enqueued on the caller's stream with stable scratch: after a
    warm-up call it can be captured in a CUDA graph and replayed on new query contents."""
    rng = np.random.default_rng(5)
    n, d, k = 20000, 256, 8
    xb = rng.standard_normal((n, d), dtype=np.float32)
    idx = pkg.B200FlatIndex(d, 0, dtype=dtype)
    idx.add(xb)
    q_static = torch.zeros((nq, d), device=cuda_device)
    q1 = torch.from_numpy(rng.standard_normal((nq, d), dtype=np.float32)).cuda()
    q2 = torch.from_numpy(rng.standard_normal((nq, d), dtype=np.float32)).cuda()
    q_static.copy_(q1)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            idx.search_ex(q_static, k)                       # warm-up: scratch reaches its final size
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = idx.search_ex(q_static, k)
    for q in (q2, q1):
        q_static.copy_(q)
        g.replay()
        torch.cuda.synchronize()
        eager = idx.search_ex(q, k)
        assert torch.equal(out["ids"], eager["ids"]) and torch.equal(out["scores"], eager["scores"])


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_very_large_query_batches_are_chunked(cuda_device, dtype):
    """More queries than one launch takes (148 SMs x 128): 74 query pairs per launch with a single bank
    split each (pacing across more pairs than a warp tracks), then a second chunk."""
    rng = np.random.default_rng(9)
    n, d, nq, k = 3000, 64, 20000, 4
    xb = rng.standard_normal((n, d), dtype=np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    idx = pkg.B200FlatIndex(d, 0, dtype=dtype)
    idx.add(xb)
    r = idx.search_ex(torch.from_numpy(xq), k)
    torch.cuda.synchronize()
    stored, qq = (xb, xq) if dtype == "fp32" else (o.bf16_round(xb), o.bf16_round(xq))
    D_ref, I_ref = o.exact_topk_f64(stored, qq, k)
    o.check_topk(stored, qq, r["scores"].cpu().numpy(), r["ids"].cpu().numpy(), 0,
                 rtol=1e-5 if dtype == "fp32" else 1e-4, D_ref=D_ref, I_ref=I_ref, what=f"chunked {dtype}")
