"""Multi-GPU row-sharded search: one process per GPU over NCCL (skipped with fewer than 2 GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import mips_oracle as o

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, nq, k, ret):
    import faulthandler

    import torch.distributed as dist

    import retrieval_augmented_mds_b200 as m

    faulthandler.dump_traceback_later(240, exit=True)     # a stuck collective must not hold the GPU box
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(5)
        xb = o.bf16_round(rng.standard_normal((n, d), dtype=np.float32))
        xq = o.bf16_round(rng.standard_normal((nq, d), dtype=np.float32))
        rows = m.shard_range(n, rank, world)  # the reference partition (mips.py:226-230)
        # facade path: build_index on this rank's rows, search with the reference call shape
        mp = m.Mips(m.MipsConfig(mips_metric_type=1, mips_normalize=True, bank_dtype="bf16"),
                    device=f"cuda:{rank}", group=dist.group.WORLD)
        mp.build_index(xb[rows.start:rows.stop])
        assert mp.index.id_offset == rows.start
        assert np.isclose(mp.phi, o.get_phi(xb), rtol=1e-6)  # all-reduce MAX over shards
        ign = rng.integers(0, n, nq).tolist()
        D, I = mp.search(mp._prepare_query(xq), ign, k)
        D_ref, I_ref = o.exact_topk_f64(xb, xq, k, ignore=np.asarray(ign))
        assert np.array_equal(np.asarray(I), I_ref)
        qn = (xq.astype(np.float64) ** 2).sum(1, keepdims=True)
        np.testing.assert_allclose(np.asarray(D), qn + o.get_phi(xb) - 2 * D_ref, rtol=1e-4, atol=1e-2)
        # device path with fused doc scores
        r = mp.search_device(torch.from_numpy(xq).cuda(), k, memory_seq_len=4)
        ids = r["ids"].cpu().numpy()
        D2, I2 = o.exact_topk_f64(xb, xq, k)
        assert np.array_equal(ids, I2)
        np.testing.assert_allclose(r["cosine"].cpu().numpy(), o.doc_scores(xq, xb[ids]), rtol=1e-4, atol=1e-5)
        assert r["memory_bias"].shape == (nq, k * 4)
        # the three spellings of the replicated-query step agree bit for bit: ONE C-ABI call with a raw
        # ncclAllGather on the step's stream ("native", the default), torch.distributed between C-ABI calls
        # ("torch"), and a CUDA-graph replay of the native step
        xq_t = torch.from_numpy(xq).cuda()
        assert mp._sharded.exchange == "native"
        sh_t = m.ShardedFlatIndex(mp.index, dist.group.WORLD, exchange="torch")
        sh_t.counts = mp._sharded.counts
        want = ("scores", "ids", "cosine", "doc_prob", "memory_bias")
        a = mp._sharded.search(xq_t, k, want=want, L=3, ignore_ids=torch.as_tensor(ign), out_mode=2)
        b = sh_t.search(xq_t, k, want=want, L=3, ignore_ids=torch.as_tensor(ign), out_mode=2)
        g = mp._sharded.capture(nq, k, with_ignore=True, want=want, L=3, out_mode=2)
        c = g.replay(xq_t, torch.as_tensor(ign).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(a["ids"].cpu().numpy(), I_ref)
        for key in ("ids", "scores", "cosine", "doc_prob", "memory_bias"):
            assert torch.equal(a[key], b[key]) and torch.equal(a[key], c[key]), key
        c2 = g.replay(xq_t.flip(0), torch.as_tensor(ign).cuda().flip(0))      # a replay with new inputs
        torch.cuda.synchronize()
        assert np.array_equal(c2["ids"].cpu().numpy(), I_ref[::-1])
        # data-parallel TRAINING step: every rank its OWN queries (retriever_generator.py:143-153 under DDP);
        # native (ncclSend/ncclRecv group) and torch (all_to_all_single) spellings agree with the oracle
        B = nq // world
        mine = slice(rank * B, (rank + 1) * B)
        for shx in (mp._sharded, sh_t):
            r = shx.search_dp(xq_t[mine], k, ignore_ids=torch.as_tensor(ign[mine]), want=("scores", "ids", "cosine"))
            torch.cuda.synchronize()
            assert np.array_equal(r["ids"].cpu().numpy(), I_ref[mine]), shx.exchange
            np.testing.assert_allclose(r["cosine"].cpu().numpy(),
                                       o.doc_scores(xq[mine], xb[r["ids"].cpu().numpy()]), rtol=1e-4, atol=1e-5)
        r = mp._sharded.search_dp(xq_t[mine], k)               # without ignored ids (no second all-gather)
        assert np.array_equal(r["ids"].cpu().numpy(), I2[mine])
        gd = mp._sharded.capture(B, k, dp=True)
        r = gd.replay(xq_t[mine])
        torch.cuda.synchronize()
        assert np.array_equal(r["ids"].cpu().numpy(), I2[mine])
        # k > 64 on the sharded bank (bounded passes per shard, lists merged by (key desc, id asc))
        Dk, Ik = o.exact_topk_f64(xb, xq, 100)
        r = mp._sharded.search(xq_t, 100, out_mode=0)
        n_diff = o.check_topk(xb, xq, r["scores"].cpu().numpy(), r["ids"].cpu().numpy(), 0, rtol=1e-4, D_ref=Dk, I_ref=Ik,
                              what="sharded k=100")           # exact up to fp32-accumulation ties among 100 results
        assert n_diff <= 0.01 * Ik.size
        r = mp._sharded.search_dp(xq_t[mine], 100, out_mode=0)
        o.check_topk(xb, xq[mine], r["scores"].cpu().numpy(), r["ids"].cpu().numpy(), 0, rtol=1e-4, D_ref=Dk[mine],
                     I_ref=Ik[mine], what="sharded dp k=100")
        # a memory refresh keeps the communicator (no new NCCL communicator / exchange buffers per rebuild)
        comm_before = mp._sharded._comm
        mp.begin_refresh(len(rows))
        mp.refresh_add(torch.from_numpy(xb[rows.start:rows.stop]).cuda())
        mp.commit_refresh(100)
        assert mp._sharded._comm is comm_before and comm_before is not None
        D3, I3 = mp.search(mp._prepare_query(xq), ign, k)
        assert np.array_equal(np.asarray(I3), I_ref)
        # peer-memory exchange (CUDA IPC over NVLink, fused into the merge kernels): same answers as the
        # NCCL all-gather, search after search (the two slot sets alternate), with and without extras
        sh = m.ShardedFlatIndex(mp.index, dist.group.WORLD, exchange="p2p")
        sh.counts = mp._sharded.counts
        for it in range(5):
            kk = k if it % 2 == 0 else 3
            a = sh.search(xq_t, kk, want=("scores", "ids", "cosine"), out_mode=0)
            b = mp._sharded.search(xq_t, kk, want=("scores", "ids", "cosine"), out_mode=0)
            torch.cuda.synchronize()
            assert torch.equal(a["ids"], b["ids"]) and torch.equal(a["scores"], b["scores"])
            assert torch.equal(a["cosine"], b["cosine"])
        a = sh.search(xq_t[:7], k, ignore_ids=torch.as_tensor(ign[:7]))
        assert np.array_equal(a["ids"].cpu().numpy(), I_ref[:7])
        sh.check_exchange()                                    # no peer wait timed out
        sh.close()
        mp._sharded.close()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_search_matches_single_bank():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(world, port, 40001, 256, 200, 8, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
