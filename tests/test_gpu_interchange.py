"""GPU tests of the drop-in routes around the index (SURVEY §8b route 1, §8f N4): the `faiss`
stand-in under an UNMODIFIED HF `datasets` (the exact call sequence of the reference,
sotasum/mips.py:333-345, :383-386, :536, :547), the flat `index.faiss` round trip and the
`Mips.save` / `Mips.load` layout."""
import io

import numpy as np
import pytest
import torch

import retrieval_augmented_mds_b200 as pkg
from oracle import mips_oracle as o

pytestmark = pytest.mark.gpu


def _data(n, d, nq, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d), dtype=np.float32), rng.standard_normal((nq, d), dtype=np.float32)


@pytest.mark.parametrize("metric", [0, 1])
def test_write_read_index_round_trip(cuda_device, metric):
    xb, xq = _data(3000, 96, 17, 3)
    idx = pkg.B200FlatIndex(96, metric, dtype="fp32")
    idx.add(xb)
    buf = io.BytesIO()
    pkg.write_index(idx, buf)
    assert buf.getvalue() == pkg.faiss_io.serialize_rows(xb, metric)      # fp32 rows are stored verbatim
    back = pkg.read_index(io.BytesIO(buf.getvalue()))
    assert (back.d, back.ntotal, back.metric_type) == (96, 3000, metric)
    D0, I0 = idx.search(xq, 5)
    D1, I1 = back.search(xq, 5)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)


def test_reference_call_sequence_through_unmodified_datasets(cuda_device, tmp_path, monkeypatch):
    """Mips.build_index / Mips.search / Mips.save / Mips.load as the reference drives HF datasets,
    with `import faiss` resolving to the stand-in."""
    pkg.install_faiss_shim()
    datasets = pytest.importorskip("datasets")
    import faiss

    if "b200" not in faiss.__version__:
        pytest.skip("a real faiss is installed")
    # this image's torchvision has no torchvision.io.VideoReader, which datasets' numpy formatter imports as
    # soon as torchvision is loaded (another test's `import transformers` loads it): unrelated to the index
    monkeypatch.setattr(datasets.config, "TORCHVISION_AVAILABLE", False)
    xb, xq = _data(2500, 64, 9, 11)
    xb_n = o.normalize_L2(xb)
    ds = datasets.Dataset.from_dict({"embeddings": xb_n})
    ds.set_format("numpy", columns=["embeddings"])
    ds.add_faiss_index(column="embeddings", index_name="mips_embeddings", string_factory="Flat",
                       train_size=None, metric_type=faiss.METRIC_INNER_PRODUCT, faiss_verbose=False)
    index = ds.get_index("mips_embeddings").faiss_index
    assert isinstance(index, pkg.B200FlatIndex) and index.ntotal == 2500   # added in 1000-row batches
    index.nprobe = 4                                                         # mips.py:342-345: accepted, ignored
    q = xq.copy()
    faiss.normalize_L2(q)                                                    # mips.py:524
    D, I = index.search(q, 6)                                                # mips.py:383-386
    D_ref, I_ref = o.exact_topk_f64(xb_n, o.normalize_L2(xq), 6)
    o.check_topk(xb_n, o.normalize_L2(xq), D, I, 0, rtol=1e-5, D_ref=D_ref, I_ref=I_ref, what="datasets route")
    scores, examples = ds.get_nearest_examples_batch("mips_embeddings", q, k=3)   # retriever_lightning.py:317-321
    assert len(examples) == 9 and np.allclose(np.asarray(scores)[:, 0], D[:, 0])
    f = tmp_path / "index.faiss"
    ds.save_faiss_index("mips_embeddings", f)                                # mips.py:536
    assert f.read_bytes()[:4] == b"IxFI"
    ds.drop_index("mips_embeddings")
    ds.load_faiss_index("mips_embeddings", f)                                # mips.py:547
    D2, I2 = ds.get_index("mips_embeddings").faiss_index.search(q, 6)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)


@pytest.mark.parametrize("metric", [0, 1])
def test_mips_facade_save_load_reference_layout(cuda_device, tmp_path, metric):
    xb, xq = _data(4000, 80, 12, 5 + metric)
    cfg = pkg.MipsConfig(mips_metric_type=metric, mips_normalize=True, mips_tmp_folder=str(tmp_path), bank_dtype="fp32")
    mips = pkg.Mips(cfg)
    mips.build_index(xb)
    s0, i0 = mips.search(xq, None, 7)
    mips.save()
    raw = (tmp_path / "mips" / "index.faiss").read_bytes()
    h = pkg.faiss_io.read_flat_header(io.BytesIO(raw).read)
    rows = np.frombuffer(raw, dtype="<f4", offset=45).reshape(h["ntotal"], h["d"])
    if metric == 0:   # IndexFlatIP over the normalised rows (mips.py:306-314)
        assert raw[:4] == b"IxFI" and h["d"] == 80
        np.testing.assert_allclose(rows, o.normalize_L2(xb), rtol=2e-7, atol=1e-8)
    else:             # IndexFlatL2 over augment_xb(rows, phi) (mips.py:316-331): d + 1 columns
        assert raw[:4] == b"IxF2" and h["d"] == 81
        want = o.augment_xb(xb, o.get_phi(xb))
        assert np.array_equal(rows[:, :80], xb)                      # L2 rows are stored raw (not normalised)
        # sqrt(phi - |x|^2): for the max-norm row the argument is 0 +- a few ulp of |x|^2 (~1e-5), i.e. ~3e-3
        np.testing.assert_allclose(rows[:, 80], want[:, 80], rtol=1e-4, atol=2e-2)
    again = pkg.Mips(cfg)
    again.load()
    assert np.isclose(again.max_norm, mips.max_norm)
    s1, i1 = again.search(xq, None, 7)
    assert np.array_equal(np.asarray(i0), np.asarray(i1))
    np.testing.assert_allclose(np.asarray(s0), np.asarray(s1), rtol=1e-5, atol=1e-4)


# ----------------------------------------------------------------------------------------- N5
def test_retriever_metrics_on_device_match_reference(cuda_device, golden):
    """Golden hit matrices generated by the reference's own retriever_metrics (pretrain.py:69-85):
    rebuild ids / labels that produce exactly that hit matrix and compute the metrics on the GPU."""
    g = golden["retriever_metrics"]
    pred, counts = g["pred"], g["counts"]
    B, k = pred.shape
    rng = np.random.default_rng(1)
    n_rows = 5000
    row_aid = rng.integers(1000, 2000, n_rows).astype(np.int64)
    query_aid = np.arange(B, dtype=np.int64)                 # labels no memory row carries ...
    ids = rng.permutation(n_rows)[: B * k].reshape(B, k).astype(np.int64)
    for b in range(B):
        row_aid[ids[b][pred[b] > 0]] = query_aid[b]          # ... except the rows that must hit
    r = pkg.retriever_metrics(torch.from_numpy(ids).cuda(), torch.from_numpy(row_aid).cuda(),
                              torch.from_numpy(query_aid).cuda(), torch.from_numpy(counts).cuda(), return_pred=True)
    assert np.array_equal(r["pred"].cpu().numpy(), pred)
    for key in ("recall", "reciprocal_rank", "average_precision"):
        assert np.isclose(r[key], float(g[key]), rtol=1e-6, atol=1e-7), key
    want = o.retriever_metrics(pred, counts)
    assert np.isclose(r["recall"], want["recall"], rtol=1e-6)


def test_retriever_metrics_edge_cases(cuda_device):
    """ids of -1 (k > ntotal padding) never hit; a first-rank hit scores reciprocal rank 0 (the
    reference's 1/argmax quirk); k up to 64."""
    ids = torch.tensor([[3, 4, -1, -1], [9, 3, 4, 5], [7, 7, 7, 7]], dtype=torch.int64).cuda()
    row_aid = torch.arange(10, dtype=torch.int64).cuda() % 5      # aid(row) = row % 5
    query_aid = torch.tensor([3, 4, 1], dtype=torch.int64).cuda()
    counts = torch.tensor([2.0, 2.0, 1.0]).cuda()
    r = pkg.retriever_metrics(ids, row_aid, query_aid, counts, return_pred=True)
    pred = r["pred"].cpu().numpy()
    assert pred.tolist() == [[1, 0, 0, 0], [1, 0, 1, 0], [0, 0, 0, 0]]
    want = o.retriever_metrics(pred, counts.cpu().numpy())
    for key in want:
        assert np.isclose(r[key], want[key], rtol=1e-6, atol=1e-7), key
    assert r["reciprocal_rank"] == 0.0
    big = torch.randint(0, 10, (5, 64), dtype=torch.int64).cuda()
    rb = pkg.retriever_metrics(big, row_aid, torch.zeros(5, dtype=torch.int64).cuda(), torch.full((5,), 3.0).cuda(),
                               return_pred=True)
    wb = o.retriever_metrics(rb["pred"].cpu().numpy(), np.full(5, 3.0, np.float32))
    for key in wb:
        assert np.isclose(rb[key], wb[key], rtol=1e-6, atol=1e-7), key
    # k > 64 (the result lists of a multi-pass search): first hit beyond rank 64 for one query
    wide = torch.randint(0, 10, (4, 150), dtype=torch.int64).cuda()
    wide[0] = 1                                                       # aid 1 everywhere ...
    wide[0, 100] = 5                                                  # ... but one row of aid 0 at rank 101
    rw = pkg.retriever_metrics(wide, row_aid, torch.zeros(4, dtype=torch.int64).cuda(), torch.full((4,), 7.0).cuda(),
                               return_pred=True)
    ww = o.retriever_metrics(rw["pred"].cpu().numpy(), np.full(4, 7.0, np.float32))
    for key in ww:
        assert np.isclose(rw[key], ww[key], rtol=1e-6, atol=1e-7), key
    assert rw["pred"][0].sum().item() == 1.0 and rw["pred"][0, 100].item() == 1.0


# ----------------------------------------------------------------------------------------- N2
def test_gather_tokens_matches_host_tokenisation(cuda_device):
    """The four tensors of mips.py:473-501 from a pre-tokenised HBM store == numpy indexing + the
    reference's mask statements (restated below), incl. -1 ids."""
    rng = np.random.default_rng(4)
    N, L, pad, bos, eos = 300, 48, 1, 0, 2
    lens = rng.integers(3, L + 1, N)
    store = np.full((N, L), pad, np.int32)
    for i, n in enumerate(lens):
        store[i, :n] = rng.integers(3, 5000, n)
        store[i, 0], store[i, n - 1] = bos, eos
    att = (np.arange(L)[None, :] < lens[:, None]).astype(np.int64)
    ms = pkg.MemoryTokenStore(store, attention_mask=att, pad_id=pad, bos_id=bos, eos_id=eos)
    ids = rng.integers(0, N, (7, 5)).astype(np.int64)
    ids[2, 3:] = -1
    out = {k: v.cpu().numpy() for k, v in ms.gather(torch.from_numpy(ids).cuda()).items()}
    flat = ids.reshape(-1)
    ok = flat >= 0
    want_ids = np.where(ok[:, None], store[np.maximum(flat, 0)], pad).astype(np.int64)
    want_att = np.where(ok[:, None], att[np.maximum(flat, 0)], 0)
    want_mem = np.where((want_ids == eos) | (want_ids == bos), 0, want_att)          # mips.py:494-501
    want_glob = np.zeros_like(want_ids)
    want_glob[:, 0] = 1                                                               # mips.py:483-487
    assert np.array_equal(out["memory_input_ids"], want_ids)
    assert np.array_equal(out["attention_mask"], want_att)
    assert np.array_equal(out["memory_attention_mask"], want_mem)
    assert np.array_equal(out["global_attention_mask"], want_glob)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_gather_rows_gives_differentiable_doc_scores(cuda_device, dtype):
    """Bank rows gathered by the search's ids reproduce the merge kernel's cosine doc scores
    (retriever_generator.py:158-172) and let the gradient reach the query."""
    xb, xq = _data(5000, 96, 6, 8)
    idx = pkg.B200FlatIndex(96, 0, dtype=dtype)
    idx.add(xb)
    q = torch.from_numpy(xq).cuda()
    r = idx.search_ex(q, 5, want=("scores", "ids", "cosine"))
    rows = idx.gather_rows(r["ids"])                                  # [6, 5, 96] fp32, no host round trip
    stored = xb if dtype == "fp32" else o.bf16_round(xb)
    assert np.array_equal(rows.cpu().numpy(), stored[r["ids"].cpu().numpy()])
    qg = (q if dtype == "fp32" else q.bfloat16().float()).clone().requires_grad_(True)
    cos = torch.nn.functional.cosine_similarity(qg[:, None, :], rows, dim=-1)
    torch.testing.assert_close(cos, r["cosine"], rtol=1e-4, atol=1e-5)
    cos.sum().backward()
    assert qg.grad is not None and bool(torch.isfinite(qg.grad).all()) and float(qg.grad.abs().sum()) > 0
    none = idx.gather_rows(torch.tensor([[-1, 10**9]], dtype=torch.int64).cuda())
    assert float(none.abs().sum()) == 0.0


# ----------------------------------------------------------------------------------------- N1
def test_double_buffered_refresh_serves_old_bank_until_commit(cuda_device):
    """Memory refresh on the device (lightning_model.py:148-180 without barriers or disk): searches
    keep seeing the old bank while the new one is ingested on a side stream; commit swaps."""
    xa, xq = _data(6000, 64, 10, 21)
    xb, _ = _data(5000, 64, 1, 22)
    cfg = pkg.MipsConfig(mips_metric_type=0, mips_normalize=True, bank_dtype="fp32")
    mips = pkg.Mips(cfg)
    mips.build_index(xa)
    xq = mips._prepare_query(xq)                                  # Mips.search takes prepared queries (mips.py:368-375)
    assert mips.needs_refresh(0, 100) is False and mips.needs_refresh(100, 100) and not mips.needs_refresh(150, 100)
    assert not mips.needs_refresh(100, 100, frozen=True)
    s_old, i_old = mips.search(xq, None, 5)
    mips.begin_refresh(5000)
    tb = torch.from_numpy(xb).cuda()
    for s in range(0, 5000, 1000):
        mips.refresh_add(tb[s:s + 1000])                          # side stream
        s_mid, i_mid = mips.search(xq, None, 5)                   # front bank, default stream
        assert np.array_equal(np.asarray(i_mid), np.asarray(i_old))
    mips.commit_refresh(100)
    assert mips.rebuilt_steps == [0, 100] and mips.index.ntotal == 5000
    s_new, i_new = mips.search(xq, None, 5)
    D_ref, I_ref = o.exact_topk_f64(o.normalize_L2(xb), xq, 5)
    o.check_topk(o.normalize_L2(xb), xq, np.asarray(s_new), np.asarray(i_new), 0, rtol=1e-5,
                 D_ref=D_ref, I_ref=I_ref, what="after refresh")
    assert np.isclose(mips.max_norm, np.sqrt(o.get_phi(xb)), rtol=1e-5)
    # the old front is the next back buffer: a second refresh reuses its allocation
    back_before = mips._back
    mips.begin_refresh(6000)
    assert mips._back is back_before and mips._back.ntotal == 0
    mips.refresh_add(torch.from_numpy(xa).cuda())
    mips.commit_refresh(200)
    s2, i2 = mips.search(xq, None, 5)
    assert np.array_equal(np.asarray(i2), np.asarray(i_old))


def test_refresh_add_waits_for_the_producer_of_the_block(cuda_device):
    """The refreshed rows come out of encoder kernels on the CALLER's stream; the ingest on the side stream must
    not read a block before its producer has written it. A long-running kernel queue on the main stream
    precedes the write of the block: without the stream dependency the back shard would ingest zeros."""
    n, d = 4096, 256
    mips = pkg.Mips(pkg.MipsConfig(mips_metric_type=0, mips_normalize=False, bank_dtype="fp32"))
    mips.build_index(np.ones((8, d), dtype=np.float32))
    mips.begin_refresh(n, d)
    blk = torch.zeros((n, d), device=cuda_device)
    big = torch.randn((8192, 8192), device=cuda_device)
    for _ in range(20):                                   # ~100 ms of queued work ahead of the block's producer
        big = big @ big
        big = big / big.abs().max()
    src = torch.randn((n, d), device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(1))
    blk.copy_(src)                                        # the "encoder" writes the block after that queue
    mips.refresh_add(blk)
    mips.commit_refresh(100)
    torch.cuda.synchronize()
    assert np.array_equal(mips.index.reconstruct_n(), src.cpu().numpy())


# ----------------------------------------------------------------------------------------- N3 (forward; gradients: test_gpu_generator_ops.py)
def test_copy_mixture_matches_reference_statements(cuda_device, golden):
    """Golden from retriever_generator.py:391-404 executed on seeded tensors, then a BART-sized vocabulary
    against the oracle and against the reference's own sequence of torch ops on the GPU."""
    g = golden["copy_mixture"]
    out = pkg.copy_mixture(*(torch.from_numpy(g[k]).cuda() for k in ("logits", "gen_gate", "copy_probs", "copy_seq")))
    np.testing.assert_allclose(out.cpu().numpy(), g["outs"], rtol=1e-5, atol=1e-5)
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    B, T, V, S = 2, 5, 50265, 2560                                   # k * L = 5 * 512 memory positions
    logits = torch.randn((B, T, V), generator=gen, device=cuda_device) * 4
    gates = torch.softmax(torch.randn((B, T, 2), generator=gen, device=cuda_device), -1)
    copy_probs = gates[..., 1:] * torch.softmax(torch.randn((B, T, S), generator=gen, device=cuda_device), -1)
    copy_seq = torch.randint(0, V, (B, S), generator=gen, device=cuda_device)
    copy_seq[:, :100] = copy_seq[:, 100:200]                         # repeated tokens accumulate
    out = pkg.copy_mixture(logits, gates[..., :1], copy_probs, copy_seq)
    probs = gates[..., :1] * torch.softmax(logits, -1)               # the reference's ops, on the GPU
    probs.scatter_add_(-1, copy_seq.reshape(B, 1, -1).expand(-1, T, -1), copy_probs)
    torch.testing.assert_close(out, torch.log(probs + 1e-7), rtol=1e-5, atol=1e-5)
    want = o.copy_mixture(logits.cpu().numpy(), gates[..., :1].cpu().numpy(), copy_probs.cpu().numpy(), copy_seq.cpu().numpy())
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=2e-5)
    # a distribution up to the reference's +1e-7 per vocabulary entry: total mass 1 + V * 1e-7
    assert bool(torch.allclose(torch.logsumexp(out, -1), torch.full((B, T), float(np.log1p(V * 1e-7)), device=cuda_device), atol=1e-4))
    with pytest.raises(pkg._lib.MipsError, match="shared memory"):
        pkg.copy_mixture(torch.zeros((1, 1, 60000), device=cuda_device), torch.ones((1, 1, 1), device=cuda_device),
                         torch.zeros((1, 1, 4), device=cuda_device), torch.zeros((1, 4), dtype=torch.int64, device=cuda_device))


def test_c_host_program_runs_through_the_abi_alone(cuda_device, tmp_path):
    """examples/c_abi_demo.c: no Python, no torch in the process — create / add / search_host from C."""
    import shutil
    import subprocess
    from pathlib import Path

    from retrieval_augmented_mds_b200 import build

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "c_abi_demo"
    subprocess.run([gcc, "-std=c99", "-I", str(root / "include"), str(root / "examples" / "c_abi_demo.c"), "-L",
                    str(build.PKG_DIR), "-lmips_b200", f"-Wl,-rpath,{build.PKG_DIR}", "-o", str(exe)], check=True)
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "query 3:  300 (" in run.stdout and "kernel: tcx" in run.stdout
