"""GPU bring-up aid for the tcgen05 search kernel: one bank tile (64 rows) and k=64 expose every
score of the 128x64 accumulator, so layout / descriptor mistakes show up as structured errors."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m
from oracle import mips_oracle as o


def full_tile(d, n=64, nq=128, seed=0, algo="tc", kmax=64):
    rng = np.random.default_rng(seed)
    xb = o.bf16_round(rng.standard_normal((n, d), dtype=np.float32))
    xq = o.bf16_round(rng.standard_normal((nq, d), dtype=np.float32))
    idx = m.B200FlatIndex(d, 0, dtype="bf16")
    idx.add(xb)
    k = min(kmax, n)
    r = idx.search_ex(torch.from_numpy(xq), k, algo=algo)
    torch.cuda.synchronize()
    ids = r["ids"].cpu().numpy()
    sc = r["scores"].cpu().numpy()
    S = xq.astype(np.float64) @ xb.astype(np.float64).T
    got = np.full((nq, n), np.nan)
    for q in range(nq):
        for j in range(k):
            if ids[q, j] >= 0:
                got[q, ids[q, j]] = sc[q, j]
    err = np.abs(got - S)
    bad = ~(err < 1e-3 * (1 + np.abs(S))) & ~np.isnan(got)
    # k=64 of n=128 rows: the 64 best must be present
    want_top = np.sort(S, axis=1)[:, ::-1][:, :k]
    got_top = np.sort(np.where(np.isnan(got), -np.inf, got), axis=1)[:, ::-1][:, :k]
    bad_top = ~(np.abs(want_top - got_top) < 1e-3 * (1 + np.abs(want_top)))
    if bad_top.any():
        print(f"  top-{k} mismatch in {bad_top.any(1).sum()} queries")
        bad[bad_top.any(1), :] = True
    print(f"[{algo}] d={d} n={n} nq={nq}: max err {np.nanmax(err):.3e}, bad {bad.sum()}/{bad.size}, "
          f"missing {np.isnan(got).sum()}", flush=True)
    if bad.any():
        qs, cs = np.nonzero(bad)
        print("  bad rows (queries):", np.unique(qs)[:20], " bad cols (bank rows):", np.unique(cs)[:20])
        # which k-chunks are being used? regress got on per-16-k partial products for one entry
        q, c = qs[0], cs[0]
        parts = (xq[q].astype(np.float64) * xb[c].astype(np.float64)).reshape(-1, 16).sum(1)
        print("  entry", (q, c), "got", got[q, c], "want", S[q, c], "partials16", np.round(parts, 2)[:16])
    return not bad.any()


ALGOS = tuple(sys.argv[1:]) or ("tc", "tc128", "tc2")
ok = True
for algo in ALGOS:
    for d in (64, 128, 256, 320, 768) + ((576, 1024) if algo == "tc2" else ()):
        kmax = 16 if d > 768 else 64
        ok &= full_tile(d, algo=algo, kmax=kmax)
        if algo == "tc2":
            ok &= full_tile(d, n=128, nq=256, algo=algo, kmax=kmax)
    ok &= full_tile(768, n=128, algo=algo)
    ok &= full_tile(192, n=100, nq=37, algo=algo)
for n in (128, 640, 6400 + 17):
    rng = np.random.default_rng(n)
    d, nq, k = 768, 300, 8
    xb = o.bf16_round(rng.standard_normal((n, d), dtype=np.float32))
    xq = o.bf16_round(rng.standard_normal((nq, d), dtype=np.float32))
    idx = m.B200FlatIndex(d, 0, dtype="bf16")
    idx.add(xb)
    b = idx.search_ex(torch.from_numpy(xq), k, algo="simt")
    for algo in ALGOS:
        a = idx.search_ex(torch.from_numpy(xq), k, algo=algo)
        torch.cuda.synchronize()
        same = (a["ids"] == b["ids"]).float().mean().item()
        print(f"n={n}: {algo} vs simt ids equal {same:.4f}, max score diff {(a['scores']-b['scores']).abs().max().item():.3e}", flush=True)
        ok &= same > 0.999
print("TC_OK" if ok else "TC_FAIL")
sys.exit(0 if ok else 1)
