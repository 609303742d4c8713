"""Host-side logic that needs no GPU: partition rules, config handling, and the N>1 exchange
(world_size-2 gloo processes on CPU) checked against the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mips_oracle as o
from retrieval_augmented_mds_b200 import Mips, MipsConfig, sharded


def test_shard_range_is_the_reference_partition():
    for n, g in ((10, 3), (100_000, 8), (2_000_000, 8), (1000, 1), (17, 4)):
        ours = [sharded.shard_range(n, r, g) for r in range(g)]
        ref = [o.shard_range(n, r, g) for r in range(g)]
        for a, b in zip(ours, ref):
            assert list(a) == [x for x in b if x < n]
        assert sum(len(r) for r in ours) == n
        assert ours[0].start == 0 and all(a.stop == b.start for a, b in zip(ours, ours[1:]))


def test_balanced_range_covers_all_rows():
    for n, g in ((10_000_000, 8), (1001, 4), (5, 8)):
        rs = [sharded.balanced_range(n, r, g) for r in range(g)]
        assert sum(len(r) for r in rs) == n
        assert max(len(r) for r in rs) - min(len(r) for r in rs) <= (n + g - 1) // g


def test_facade_rejects_approximate_factories_and_keeps_reference_fields():
    with pytest.raises(ValueError):
        Mips(MipsConfig(mips_string_factory="IVF256,SQ8"))
    mp_ = Mips(MipsConfig(mips_metric_type=1, mips_normalize=True, mips_tmp_folder="/tmp/x"))
    assert mp_.metric_type == 1 and mp_.normalize and mp_.string_factory == "Flat"
    assert mp_.index_name == "mips_embeddings" and mp_.embeddings_column == "embeddings"
    assert mp_.rebuilt_steps == [0] and str(mp_.max_norm_file).endswith("mips/max_norm.pkl")
    with pytest.raises(RuntimeError):
        mp_.search(np.zeros((1, 4), dtype=np.float32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        xb = rng.standard_normal((n, d), dtype=np.float32)
        xq = rng.standard_normal((nq, d), dtype=np.float32)
        rows = sharded.shard_range(n, rank, world)
        off, counts = sharded.exchange_offsets(len(rows), None, None)
        assert off == rows.start and counts == [len(sharded.shard_range(n, r, world)) for r in range(world)]
        phi = sharded.allreduce_max(float((xb[rows.start:rows.stop] ** 2).sum(1).max()))
        assert np.isclose(phi, o.get_phi(xb), rtol=1e-6)
        # the local search is CUDA-only; stand in for it with the oracle so that the exchange
        # and layout logic can be exercised on CPU
        D, I = o.flat_search(xb[rows.start:rows.stop], xq, k)
        xn2 = (xb[rows.start:rows.stop][I] ** 2).sum(-1).astype(np.float32)
        g_key, g_ids, g_xn2 = sharded.gather_candidates(torch.from_numpy(D), torch.from_numpy(I + off),
                                                        torch.from_numpy(xn2))
        assert g_key.shape == (world, nq, k) and g_ids.dtype == torch.int64
        assert torch.equal(g_key[rank], torch.from_numpy(D)) and torch.equal(g_ids[rank], torch.from_numpy(I + off))
        assert torch.equal(g_xn2[rank], torch.from_numpy(xn2))
        # the packed exchange used by the product path: {f32 key, f32 |x|^2, i64 id} records
        rec = np.zeros((nq, k), dtype=np.dtype([("key", "<f4"), ("xn2", "<f4"), ("id", "<i8")]))
        rec["key"], rec["xn2"], rec["id"] = D, xn2, I + off
        assert rec.dtype.itemsize == 16
        packed = torch.from_numpy(rec.view(np.uint8).reshape(nq, k, 16).copy())
        g_packed = sharded.gather_packed(packed)
        assert g_packed.shape == (world, nq, k, 16)
        back = g_packed.numpy().reshape(world, nq, k * 16).view(rec.dtype).reshape(world, nq, k)
        assert np.array_equal(back["key"], g_key.numpy()) and np.array_equal(back["id"], g_ids.numpy())
        # merging the gathered lists (oracle ordering rule) reproduces the unsharded search
        cat_s = g_key.permute(1, 0, 2).reshape(nq, -1).numpy()
        cat_i = g_ids.permute(1, 0, 2).reshape(nq, -1).numpy()
        order = np.lexsort((cat_i, -cat_s), axis=1)[:, :k]
        D_ref, I_ref = o.flat_search(xb, xq, k)
        assert np.array_equal(np.take_along_axis(cat_i, order, 1), I_ref)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def _dp_worker(rank, world, port, n, d, B, k, ret):
    """Exchange logic of the data-parallel search (every rank its own B queries): all-gather of the queries,
    local search of all G*B (the oracle stands in for the CUDA-only K1), all-to-all of the 16-byte records,
    merge of the rank's own B queries == unsharded search of those queries."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(17)
        xb = rng.standard_normal((n, d), dtype=np.float32)
        xq_all = rng.standard_normal((world * B, d), dtype=np.float32)
        mine = xq_all[rank * B:(rank + 1) * B]
        rows = sharded.shard_range(n, rank, world)
        q_all = torch.empty((world * B, d), dtype=torch.float32)
        dist.all_gather_into_tensor(q_all, torch.from_numpy(mine.copy()))
        assert np.array_equal(q_all.numpy(), xq_all)
        D, I = o.flat_search(xb[rows.start:rows.stop], q_all.numpy(), k)
        rec = np.zeros((world * B, k), dtype=np.dtype([("key", "<f4"), ("xn2", "<f4"), ("id", "<i8")]))
        rec["key"], rec["id"] = D, np.where(I >= 0, I + rows.start, -1)
        packed = torch.from_numpy(rec.view(np.uint8).reshape(world * B, k, 16).copy())
        got = sharded.alltoall_packed(packed)
        assert got.shape == (world, B, k, 16)
        back = got.numpy().reshape(world, B, k * 16).view(rec.dtype).reshape(world, B, k)
        cat_s = back["key"].transpose(1, 0, 2).reshape(B, -1)
        cat_i = back["id"].transpose(1, 0, 2).reshape(B, -1)
        order = np.lexsort((cat_i, -cat_s), axis=1)[:, :k]
        D_ref, I_ref = o.flat_search(xb, mine, k)
        assert np.array_equal(np.take_along_axis(cat_i, order, 1), I_ref)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_exchange_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_dp_worker, args=(world, port, 3001, 24, 5, 4, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_two_rank_exchange_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(world, port, 5003, 32, 9, 6, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_weighted_ranges_tile_the_bank_and_follow_the_weights():
    from retrieval_augmented_mds_b200.sharded import weighted_ranges

    for n, w in ((10_000_000, [1.0, 0.94, 1.03, 0.97, 1.0, 1.06, 0.99, 1.01]), (1001, [1, 1, 1]), (5, [3, 1]), (0, [1, 2])):
        rs = weighted_ranges(n, w)
        assert len(rs) == len(w) and rs[0].start == 0 and rs[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(rs, rs[1:]))
    rs = weighted_ranges(8_000_000, [1.0, 0.5, 1.0, 2.0])          # clamped to +-10 % of the equal share
    sizes = [len(r) for r in rs]
    assert max(sizes) <= 1.2 * 2_000_000 and min(sizes) >= 0.8 * 2_000_000 and sizes[3] > sizes[0] == sizes[1]
    rs = weighted_ranges(8_000_000, [1.00, 0.95, 1.05, 1.00])
    sizes = [len(r) for r in rs]
    assert sizes[2] > sizes[0] > sizes[1] and abs(sizes[2] / sizes[1] - 1.05 / 0.95) < 1e-3
    with pytest.raises(ValueError):
        weighted_ranges(10, [0, 0])


def test_finalize_lists_matches_merge_kernel_contract(golden):
    """finalize_lists = the output transform of the merge kernel spelled with tensor ops (used for k > 64, where the
    lists are longer than one kernel pass holds): against the reference's doc-score statements (golden) and the
    oracle's metric transforms, on CPU tensors."""
    from retrieval_augmented_mds_b200.index import finalize_lists
    g = golden["doc_scores"]
    q, docs = torch.from_numpy(g["query"]), torch.from_numpy(g["docs"])            # [B, d], [B, k, d]
    B, k, _ = docs.shape
    L = int(g["memory_seq_len"])
    ip = (q[:, None, :] * docs).sum(-1)
    xn2, qn2 = (docs ** 2).sum(-1), (q ** 2).sum(-1)
    ids = torch.arange(B * k).view(B, k)
    out = finalize_lists(ip, ids, xn2, qn2, 0, None, 0.0, ("scores", "ids", "cosine", "doc_prob", "memory_bias"), L, 1.5, -0.25)
    np.testing.assert_allclose(out["scores"].numpy(), ip.numpy(), rtol=1e-6)
    np.testing.assert_allclose(out["cosine"].numpy(), g["mips_scores"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["memory_bias"].numpy(), g["memory_bias"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["doc_prob"].numpy(), o.doc_prob(g["mips_scores"], 1.5, -0.25), rtol=1e-5, atol=1e-6)
    # L2 metric: the ranking key is <q,x> - |x|^2/2; IndexFlatL2 distances, and the augmented form |q|^2 + phi - 2<q,x>
    key_l2 = ip - 0.5 * xn2
    d_l2 = finalize_lists(key_l2, ids, xn2, qn2, 1, 1, 0.0, ("scores", "ids"), None, 1.0, 0.0)["scores"]
    np.testing.assert_allclose(d_l2.numpy(), ((q[:, None, :] - docs) ** 2).sum(-1).numpy(), rtol=1e-4, atol=1e-4)
    phi = float(xn2.max())
    d_aug = finalize_lists(ip, ids, xn2, qn2, 0, 2, phi, ("scores", "ids"), None, 1.0, 0.0)["scores"]
    np.testing.assert_allclose(d_aug.numpy(), (qn2[:, None] + phi - 2 * ip).numpy(), rtol=1e-5, atol=1e-5)
    # padding (ids -1): -inf / +inf scores, zero cosine, zero probability mass
    ids_pad = ids.clone()
    ids_pad[:, -2:] = -1
    pad = finalize_lists(ip, ids_pad, xn2, qn2, 0, None, 0.0, ("scores", "ids", "cosine", "doc_prob"), None, 1.0, 0.0)
    assert torch.isinf(pad["scores"][:, -2:]).all() and (pad["cosine"][:, -2:] == 0).all()
    assert (pad["doc_prob"][:, -2:] == 0).all()
    np.testing.assert_allclose(pad["doc_prob"].sum(1).numpy(), 1.0, rtol=1e-5)


def test_bench_reference_arm_runs_real_steps_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours): on a small bank it must run REAL
    steps — value = queries of a step / measured step time — and print one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, BENCH_REF_BUDGET_S="6", OMP_NUM_THREADS="4")
    res = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--rows", "150000", "--nq", "64",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mips_queries_per_s" and d["unit"] == "queries/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    nq_s = d["config"]["queries_per_step"]
    assert abs(d["value"] - nq_s / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]          # nothing extrapolated
    assert d["ms_per_step"] * d["steps"] * 1e-3 <= d["wall_s"]                              # the steps fit the run


class _OracleShard:
    """Stand-in for B200FlatIndex on CPU (the local search is CUDA only): the oracle searches this rank's rows, so
    that the host-side exchange / merge logic of ShardedFlatIndex can run under gloo."""

    dtype, metric_type, phi = "fp32", 0, 0.0

    def __init__(self, rows: np.ndarray, id_offset: int):
        self.rows, self.id_offset = rows, id_offset
        self.d = rows.shape[1]
        self.device = torch.device("cpu")

    def _check_k(self, k, multipass=False):
        return int(k)

    def search_local_multipass(self, xq, k, ignore_ids=None, normalize_queries=False, algo="auto"):
        q = xq.numpy()
        ign = None if ignore_ids is None else ignore_ids.numpy() - self.id_offset
        D, I = o.exact_topk_f64(self.rows, q, k, 0, ignore=ign)
        if I.shape[1] < k:                                    # the index pads with id -1 when k exceeds its rows
            pad = k - I.shape[1]
            D = np.concatenate([D, np.full((len(q), pad), -np.inf)], 1)
            I = np.concatenate([I, np.full((len(q), pad), -1, dtype=np.int64)], 1)
        xn2 = np.where(I >= 0, (self.rows[np.maximum(I, 0)] ** 2).sum(-1), 0.0)
        ids = np.where(I >= 0, I + self.id_offset, -1)
        key = np.where(I >= 0, D, -np.inf)
        return (torch.from_numpy(key.astype(np.float32)), torch.from_numpy(ids.astype(np.int64)),
                torch.from_numpy(xn2.astype(np.float32)), torch.from_numpy((q ** 2).sum(1).astype(np.float32)))


def _bigk_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(23)
        n, d, nq, k = 700, 24, 12, 100
        xb = rng.standard_normal((n, d)).astype(np.float32)
        xb[300:305] = xb[3]                                   # ties across the two shards
        xq = rng.standard_normal((nq, d)).astype(np.float32)
        xq[0] = xb[3]
        rows = sharded.shard_range(n, rank, world)
        sh = sharded.ShardedFlatIndex(_OracleShard(xb[rows.start:rows.stop], rows.start), exchange="torch")
        ign = rng.integers(0, n, nq)
        D_ref, I_ref = o.exact_topk_f64(xb, xq, k, 0, ignore=ign)
        r = sh.search(torch.from_numpy(xq), k, ignore_ids=torch.from_numpy(ign), want=("scores", "ids", "cosine"))
        assert np.array_equal(r["ids"].numpy(), I_ref)        # (key desc, id asc) across shards, k > 64
        np.testing.assert_allclose(r["scores"].numpy(), D_ref, rtol=1e-5, atol=1e-5)
        B = nq // world
        mine = slice(rank * B, (rank + 1) * B)
        r = sh.search_dp(torch.from_numpy(xq[mine]), k, ignore_ids=torch.from_numpy(ign[mine]))
        assert np.array_equal(r["ids"].numpy(), I_ref[mine])  # every rank its own queries
        r = sh.search(torch.from_numpy(xq), 690)              # k beyond each shard's rows: -1 padding merges away
        assert np.array_equal(r["ids"].numpy(), o.exact_topk_f64(xb, xq, 690)[1])
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_large_k_merge_gloo():
    """k > 64 on a row-sharded bank: per-shard lists (here from the oracle) gathered through torch.distributed and
    merged by (key descending, id ascending) — replicated and data-parallel queries, ties across shards."""
    world, port = 2, _free_port()
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_bigk_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
