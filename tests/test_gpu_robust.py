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


def test_full_size_bank_properties(cuda_device):
    """BASELINE config 3 at its FULL size (10M x 768 bf16, 1024 queries, k=8) through size-independent
    properties: planted neighbours come first, scores descend, ids are distinct and valid, and the
    answer equals the merge of two half banks searched separately (top-k of a union = merge of the
    parts' top-k: the identity the row-sharded search relies on)."""
    n, d, nq, k = 10_000_000, 768, 1024, 8
    gen = torch.Generator(device=cuda_device).manual_seed(2024)
    full = pkg.B200FlatIndex(d, 0, dtype="bf16", capacity=n)
    lo = pkg.B200FlatIndex(d, 0, dtype="bf16", capacity=n // 2)
    hi = pkg.B200FlatIndex(d, 0, dtype="bf16", capacity=n // 2, id_offset=n // 2)
    planted_rows = torch.randint(0, n, (nq,), generator=gen, device=cuda_device)
    xq = torch.empty((nq, d), device=cuda_device)
    chunk = 500_000
    for s in range(0, n, chunk):
        blk = torch.randn((chunk, d), generator=gen, device=cuda_device)
        sel = (planted_rows >= s) & (planted_rows < s + chunk)
        if bool(sel.any()):
            xq[sel] = blk[planted_rows[sel] - s].bfloat16().float() * 1.5   # the planted row wins by a wide margin
        full.add(blk)
        (lo if s < n // 2 else hi).add(blk)
    r = full.search_ex(xq, k)
    ids, sc = r["ids"], r["scores"]
    assert full.last_algo == "tc2"
    assert torch.equal(ids[:, 0], planted_rows)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all()) and int(ids.min()) >= 0 and int(ids.max()) < n
    assert all(len(set(row)) == k for row in ids[:64].cpu().tolist())
    p_lo, qn2 = lo.search_local_packed(xq, k)
    p_hi, _ = hi.search_local_packed(xq, k)
    merged = full.merge_packed(torch.stack([p_lo, p_hi]), qn2, k)
    assert torch.equal(merged["ids"], ids)
    torch.testing.assert_close(merged["scores"], sc, rtol=1e-5, atol=1e-3)
