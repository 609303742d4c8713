"""Parity of the CUDA path (through the C ABI) with the oracle and the golden fixtures.

Bars (BASELINE.json north_star): fp32 bank — ids identical to exact search except ties inside
rtol=1e-5, scores within 1e-5 relative; bf16 bank — recall@k = 1.0 against fp32 flat IP ON THE
SAME (bf16-rounded) INPUTS, scores within 1e-2 relative (we hold 1e-4: products of bf16 values
are exact in fp32, only the accumulation order differs).
"""
import numpy as np
import pytest
import torch

from oracle import mips_oracle as o

pytestmark = pytest.mark.gpu

RTOL_F32 = 1e-5
RTOL_BF16 = 1e-4


@pytest.fixture(scope="module")
def m(cuda_device):
    import retrieval_augmented_mds_b200 as pkg
    return pkg


def _data(n, d, nq, seed=0, scale_rows=True):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d), dtype=np.float32)
    if scale_rows:
        xb *= rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    return xb, xq


# ----------------------------------------------------------------------------------------- K0
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("d", [64, 96, 768, 100])
def test_add_stores_rows_norms_and_max(m, dtype, d):
    xb, _ = _data(1000, d, 1, seed=d)
    idx = m.B200FlatIndex(d, m.METRIC_INNER_PRODUCT, dtype=dtype)
    idx.add(xb[:300])
    idx.add(torch.from_numpy(xb[300:]).cuda())  # device path, appended
    assert idx.ntotal == 1000
    rows = idx.reconstruct_n(0, 1000)
    want = xb if dtype == "fp32" else o.bf16_round(xb)
    assert np.array_equal(rows, want)  # bit exact: fp32 copy / RNE bf16 rounding
    assert np.isclose(idx.max_norm2(), o.get_phi(xb), rtol=1e-5)
    idx.reset()
    assert idx.ntotal == 0


def test_add_normalize_matches_faiss_contract(m):
    xb, _ = _data(500, 96, 1, seed=5)
    xb[7] = 0.0
    idx = m.B200FlatIndex(96, m.METRIC_INNER_PRODUCT, dtype="fp32")
    idx.add(xb, normalize=True)
    np.testing.assert_allclose(idx.reconstruct_n(), o.normalize_L2(xb), rtol=2e-7, atol=1e-8)
    assert np.isclose(idx.max_norm2(), o.get_phi(xb), rtol=1e-5)  # max over the RAW rows (mips.py:298-304)
    x = xb.copy()
    m.normalize_L2(x)  # faiss.normalize_L2 shim, in place
    np.testing.assert_allclose(x, o.normalize_L2(xb), rtol=2e-7, atol=1e-8)
    assert np.all(x[7] == 0)


# ----------------------------------------------------------------------------------------- K1+K2
def _search_np(idx, xq, k, **kw):
    r = idx.search_ex(torch.from_numpy(xq), k, **kw)
    torch.cuda.synchronize()
    return {key: v.cpu().numpy() for key, v in r.items()}


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,nq,k", [(5000, 96, 37, 8), (20000, 768, 256, 8), (3001, 64, 1, 1),
                                      (777, 200, 130, 33), (4096, 128, 64, 64)])
def test_fp32_search_is_exact(m, metric, n, d, nq, k):
    xb, xq = _data(n, d, nq, seed=n + k)
    idx = m.B200FlatIndex(d, metric, dtype="fp32")
    idx.add(xb)
    r = _search_np(idx, xq, k)
    # AUTO on an fp32 bank: tensor-core filter + exact re-rank + certificate (SIMT only as its fallback)
    assert idx.last_algo == "tcx"
    n_diff = o.check_topk(xb, xq, r["scores"], r["ids"], metric, rtol=RTOL_F32, what=f"fp32 m{metric}")
    assert n_diff <= max(1, nq * k // 500)
    D, I = idx.search(xq, k)  # numpy in/out = the host end-to-end entry point
    assert np.array_equal(I, r["ids"]) and np.array_equal(D, r["scores"])
    # the plain fp32-FMA kernel (the fallback of the certificate) gives the same answer
    s = _search_np(idx, xq, k, algo="simt")
    assert idx.last_algo == "simt"
    o.check_topk(xb, xq, s["scores"], s["ids"], metric, rtol=RTOL_F32, what=f"fp32 simt m{metric}")
    assert (s["ids"] != r["ids"]).mean() <= 2e-3  # only ties inside the fp32 rounding tolerance may differ
    np.testing.assert_allclose(s["scores"], r["scores"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("metric", [0, 1])
def test_fp32_exact_certificate_and_fallback(m, metric):
    """Exact tensor-core search on an fp32 bank (MIPS_ALGO_TCX): random data certifies without any
    fallback; exact duplicates straddling the k-th place cannot be certified and go through the
    SIMT kernel — both give the exact answer with the (score desc, id asc) tie rule."""
    n, d, nq, k = 30000, 256, 300, 8
    xb, xq = _data(n, d, nq, seed=77, scale_rows=False)
    idx = m.B200FlatIndex(d, metric, dtype="fp32")
    idx.add(xb)
    idx.fallback_queries(reset=True)
    r = _search_np(idx, xq, k, algo="tcx")
    assert idx.last_algo == "tcx"
    assert idx.fallback_queries() == 0
    o.check_topk(xb, xq, r["scores"], r["ids"], metric, rtol=RTOL_F32, what="tcx random")
    # 40 copies of one row: for queries aimed at it the k-th and the kc-th candidates tie exactly
    xb2 = xb.copy()
    dup = np.arange(100, 100 + 40 * 700, 700)
    xb2[dup] = xb2[100]
    xq2 = xq.copy()
    xq2[:50] = xb2[100] * 3 + 0.01 * xq[:50]
    idx2 = m.B200FlatIndex(d, metric, dtype="fp32")
    idx2.add(xb2)
    idx2.fallback_queries(reset=True)
    r2 = _search_np(idx2, xq2, k, algo="tcx")
    n_fb = idx2.fallback_queries()
    assert 50 <= n_fb <= nq
    for q in range(50):
        assert r2["ids"][q].tolist() == dup[:k].tolist()  # lowest ids among the exact ties
    o.check_topk(xb2, xq2, r2["scores"], r2["ids"], metric, rtol=RTOL_F32, what="tcx ties")
    s2 = _search_np(idx2, xq2, k, algo="simt")
    assert np.array_equal(s2["ids"][:50], r2["ids"][:50])


@pytest.mark.parametrize("algo", ["tc", "tc128", "tc2", "simt"])
@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,nq,k", [(5000, 96, 37, 8), (20000, 768, 256, 8), (3001, 64, 1, 1),
                                      (777, 256, 130, 33), (4096, 128, 64, 64), (64 * 300 + 5, 768, 300, 5),
                                      (10000, 320, 129, 16)])
def test_bf16_search_matches_fp32_on_same_inputs(m, algo, metric, n, d, nq, k):
    xb, xq = _data(n, d, nq, seed=n + k + 1)
    idx = m.B200FlatIndex(d, metric, dtype="bf16")
    idx.add(xb)
    r = _search_np(idx, xq, k, algo=algo)
    assert idx.last_algo == algo
    xb_r, xq_r = o.bf16_round(xb), o.bf16_round(xq)  # "same inputs": the rounded values, in fp32
    D_ref, I_ref = o.exact_topk_f64(xb_r, xq_r, k, metric)
    o.check_topk(xb_r, xq_r, r["scores"], r["ids"], metric, rtol=RTOL_BF16, D_ref=D_ref, I_ref=I_ref,
                 what=f"bf16 {algo} m{metric}")
    assert o.recall_at_k(I_ref, r["ids"]) >= 1.0 - 1e-3  # 1.0 up to certified ties (checked above)


def test_tc_and_simt_agree_bit_for_bit_on_ids(m):
    xb, xq = _data(50000, 768, 200, seed=11)
    idx = m.B200FlatIndex(768, 0, dtype="bf16")
    idx.add(xb)
    a = _search_np(idx, xq, 10, algo="tc")
    b = _search_np(idx, xq, 10, algo="simt")
    assert (a["ids"] != b["ids"]).mean() < 1e-3
    np.testing.assert_allclose(a["scores"], b["scores"], rtol=1e-5, atol=1e-4)


def test_golden_inner_product_through_cuda(m, golden):
    g, inp = golden["inner_product"], golden["inputs"]
    xb, xq = inp["xb"], inp["xq"]
    idx = m.B200FlatIndex(xb.shape[1], 0, dtype="fp32")
    idx.add(xb)
    for k in (1, 8, 10):
        D, I = idx.search(xq, k)
        assert np.array_equal(I, g[f"ids_n0_k{k}"])
        np.testing.assert_allclose(D, g[f"scores_n0_k{k}"], rtol=1e-5, atol=1e-5)
    # normalize=True (the reference default): unit rows + unit queries
    idn = m.B200FlatIndex(xb.shape[1], 0, dtype="fp32")
    idn.add(xb, normalize=True)
    r = _search_np(idn, xq, 10, normalize_queries=True)
    assert np.array_equal(r["ids"], g["ids_n1_k10"])
    np.testing.assert_allclose(r["scores"], g["scores_n1_k10"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("bank_dtype", ["fp32"])
def test_np_search_facade_matches_reference_inner_product(m, golden, bank_dtype):
    """`Mips.np_search` (mips.py:527-529 -> inner_product :552-560) against the reference's own outputs, for the
    three (metric, normalize) combinations the config allows — including normalize with the L2 metric, where the
    stored rows are not unit-norm and the reference's per-call renormalisation ranks by cosine."""
    g, inp = golden["inner_product"], golden["inputs"]
    xb, xq = inp["xb"], inp["xq"]
    for metric, normalize, tag in ((0, False, "n0"), (0, True, "n1"), (1, True, "n1"), (1, False, "n0")):
        mp = m.Mips(m.MipsConfig(mips_metric_type=metric, mips_normalize=normalize, bank_dtype=bank_dtype))
        mp.build_index(xb)
        for k in (1, 8, 10):
            S, I = mp.np_search(xq, k)
            assert S.dtype == np.float32 and I.dtype == np.int64 and S.shape == (len(xq), k)
            assert np.array_equal(I, g[f"ids_{tag}_k{k}"]), (metric, normalize, k)
            np.testing.assert_allclose(S, g[f"scores_{tag}_k{k}"], rtol=1e-5, atol=1e-5)
        if metric == 1 and normalize:       # the unit-norm copy is built once per bank, not per call
            assert mp._unit_bank.ntotal == len(xb) and mp._unit_bank_key == (id(mp.index), len(xb))


def test_result_containers_carry_the_reference_field_names(m, golden):
    """A12: MipsModelOutput (mips.py:33-42) and RGEncoderModelOutput (retriever_generator.py:29-42) — `scores` /
    `faiss_scores` are the raw search scores, `mips_scores` the cosine re-score (golden from the reference's
    statements), `memory_bias` its broadcast over each document's tokens, `query_cls` the queries."""
    g = golden["doc_scores"]
    query, docs = g["query"], g["docs"]                     # [B, d], [B, k, d]: every query has its own k documents
    B, k, d = docs.shape
    L = int(g["memory_seq_len"])
    bank = docs.reshape(B * k, d)
    mp = m.Mips(m.MipsConfig(mips_metric_type=0, mips_normalize=False, bank_dtype="fp32"))
    mp.build_index(bank)
    # token store: document r has r % L + 1 tokens, ids 100 + r
    toks = np.full((B * k, L), 1, dtype=np.int64)
    lens = np.array([r % L + 1 for r in range(B * k)])
    for r in range(B * k):
        toks[r, :lens[r]] = 100 + r
    store = m.MemoryTokenStore(toks, lengths=lens, pad_id=1, bos_id=0, eos_id=2)
    q_dev = torch.from_numpy(query).cuda()
    out = mp.retrieve_for_generator(q_dev[:, None, :], B * k, token_store=store)      # all documents, ranked
    assert set(vars(out)) >= {"mips_scores", "faiss_scores", "memory_bias", "query_cls", "memory_mask", "copy_sequence",
                              "examples"}
    ids = out.examples.cpu().numpy()
    ip = query @ bank.T
    np.testing.assert_allclose(out.faiss_scores.cpu().numpy(), np.take_along_axis(ip, ids, 1), rtol=1e-5, atol=1e-5)
    # the reference's cosine of query b with its own documents, looked up through the returned ids
    own = np.stack([[np.where(ids[b] == b * k + j)[0][0] for j in range(k)] for b in range(B)])
    got = np.take_along_axis(out.mips_scores.cpu().numpy(), own, 1)
    np.testing.assert_allclose(got, g["mips_scores"], rtol=1e-5, atol=1e-6)
    mb = out.memory_bias.cpu().numpy().reshape(B, B * k, L)
    assert np.array_equal(mb, np.repeat(out.mips_scores.cpu().numpy()[:, :, None], L, axis=2))
    assert torch.equal(out.query_cls, q_dev)
    assert out.copy_sequence.shape == (B, B * k * L) and out.memory_mask.shape == (B, B * k * L)
    assert torch.equal(out.copy_sequence.view(B, B * k, L)[0, 0], torch.from_numpy(toks[ids[0, 0]]).cuda())
    # Mips.forward: numpy queries like the reference, MipsModelOutput field names, metrics on request
    aid_rows = np.arange(B * k) // k                          # documents of query b share aid b
    fo = mp.forward(query, k=k, aid=np.arange(B), aid_counts=np.full(B, k, np.float32), row_aid=aid_rows,
                    token_store=store)
    assert set(vars(fo)) >= {"scores", "examples", "query_cls", "metrics", "memory_input_ids", "memory_attention_mask"}
    assert isinstance(fo.query_cls, np.ndarray) and np.array_equal(fo.query_cls, query)
    np.testing.assert_allclose(fo.scores.cpu().numpy(), np.sort(ip, 1)[:, ::-1][:, :k], rtol=1e-5, atol=1e-5)
    assert fo.memory_input_ids.shape == (B, k, L) and set(fo.metrics) == {"recall", "reciprocal_rank", "average_precision"}


def test_golden_mips_search_and_ignore_through_facade(m, golden):
    g, inp = golden["mips_search"], golden["inputs"]
    xb, xq = inp["xb"], inp["xq"]
    mp = m.Mips(m.MipsConfig(mips_metric_type=0, mips_normalize=False, bank_dtype="fp32"))
    mp.build_index(xb)
    s0, i0 = mp.search(mp._prepare_query(xq), None, 10)
    assert isinstance(s0, np.ndarray) and np.array_equal(i0, g["ids_plain"])
    np.testing.assert_allclose(s0, g["scores_plain"], rtol=1e-5, atol=1e-5)
    s1, i1 = mp.search(mp._prepare_query(xq), g["ignore"].tolist(), 10)
    assert isinstance(s1, list) and isinstance(i1, list) and isinstance(i1[0], list)
    assert np.array_equal(np.asarray(i1), g["ids_ignore"])
    np.testing.assert_allclose(np.asarray(s1), g["scores_ignore"], rtol=1e-5, atol=1e-5)


def test_golden_augmented_l2_equals_ip(m, golden):
    """test_faiss_index invariant (mips.py:655-685) + distances of IndexFlatL2 on augmented vectors."""
    g, inp = golden["augment"], golden["inputs"]
    xb, xq = inp["xb"], inp["xq"][:8]
    mp = m.Mips(m.MipsConfig(mips_metric_type=1, mips_normalize=True, bank_dtype="fp32"))
    mp.build_index(xb)
    assert np.isclose(mp.phi, float(g["phi"]), rtol=1e-6)
    q = mp._prepare_query(xq)
    assert q.shape[1] == xb.shape[1] + 1 and np.array_equal(q, g["xq_aug"])
    D, I = mp.search(q, None, 10)
    assert np.array_equal(I, g["ids_l2"]) and np.array_equal(I, g["ids_ip"])
    np.testing.assert_allclose(D, g["d2"], rtol=1e-4, atol=2e-3)
    # the same through a true L2 index on physically augmented rows (faiss route)
    aug = o.augment_xb(xb)
    idx = m.IndexFlatL2(aug.shape[1], dtype="fp32")
    idx.add(aug)
    D2, I2 = idx.search(g["xq_aug"], 10)
    assert np.array_equal(I2, g["ids_l2"])
    np.testing.assert_allclose(D2, g["d2"], rtol=1e-4, atol=2e-3)


def test_golden_prepare_query(m, golden):
    g = golden["prepare_query"]
    for metric, norm in ((0, True), (0, False), (1, True)):
        mp = m.Mips(m.MipsConfig(mips_metric_type=metric, mips_normalize=norm))
        out = mp._prepare_query(g["xq"].copy())
        ref = g[f"m{metric}_n{int(norm)}"]
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == ref.shape
        np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-7)


def test_golden_doc_scores_fused_in_merge(m, golden):
    """retriever_generator.py:158-172,188-192 when the doc encoder is frozen: the re-encoded CLS
    equals the stored row, so cosine/memory_bias come straight out of K2."""
    g = golden["doc_scores"]
    query, docs, L = g["query"], g["docs"], int(g["memory_seq_len"])
    B, K, d = docs.shape
    bank = docs.reshape(B * K, d)
    idx = m.B200FlatIndex(d, 0, dtype="fp32")
    idx.add(bank)
    r = _search_np(idx, query, B * K, want=("scores", "ids", "cosine"))
    for b in range(B):
        # pick the cosine of this query's own K docs out of the full ranking
        pos = {int(i): j for j, i in enumerate(r["ids"][b])}
        cos = np.array([r["cosine"][b, pos[b * K + j]] for j in range(K)])
        np.testing.assert_allclose(cos, g["mips_scores"][b], rtol=1e-5, atol=1e-6)
    # memory_bias / doc_prob layout on a plain top-K search
    r = _search_np(idx, query, K, want=("scores", "ids", "cosine", "doc_prob", "memory_bias"), L=L)
    np.testing.assert_array_equal(r["memory_bias"], o.memory_bias(r["cosine"], L))
    np.testing.assert_allclose(r["doc_prob"], o.doc_prob(r["cosine"]), rtol=1e-5, atol=1e-6)
    rows = bank[r["ids"]]
    np.testing.assert_allclose(r["cosine"], o.doc_scores(query, rows), rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_edge_cases(m, dtype):
    d = 64
    idx = m.B200FlatIndex(d, 0, dtype=dtype)
    xq = np.ones((3, d), dtype=np.float32)
    D, I = idx.search(xq, 4)  # empty index: faiss returns -1 ids
    assert (I == -1).all() and np.isneginf(D).all()
    xb, _ = _data(5, d, 1, seed=2)
    idx.add(xb)
    D, I = idx.search(xq, 8)  # k > ntotal: -1 / -inf padding
    assert (I[:, 5:] == -1).all() and np.isneginf(D[:, 5:]).all()
    assert sorted(I[0, :5].tolist()) == [0, 1, 2, 3, 4]
    D0, I0 = idx.search(np.zeros((0, d), dtype=np.float32), 3)  # empty query batch
    assert D0.shape == (0, 3) and I0.shape == (0, 3)
    with pytest.raises(ValueError):
        idx.search(np.ones((2, d + 1), dtype=np.float32), 3)
    with pytest.raises(ValueError):
        idx.search(np.ones((d,), dtype=np.float32), 3)
    with pytest.raises(ValueError):
        idx.search(xq, 0)
    with pytest.raises(ValueError):
        idx.add(np.ones((2, d + 3), dtype=np.float32))
    with pytest.raises(ValueError):
        m.index_factory(d, "IVF256,SQ8", 0)
    l2 = m.B200FlatIndex(d, 1, dtype=dtype)
    l2.add(xb)
    D, I = l2.search(xq, 8)
    assert (I[:, 5:] == -1).all() and np.isposinf(D[:, 5:]).all()


@pytest.mark.parametrize("dtype,algo", [("fp32", "simt"), ("fp32", "tcx"), ("bf16", "tc"), ("bf16", "tc2"), ("bf16", "simt")])
def test_duplicates_and_ties_resolve_to_lower_id(m, dtype, algo):
    d, n = 128, 3000
    xb, xq = _data(n, d, 20, seed=9, scale_rows=False)
    xb = o.bf16_round(xb)
    xq = o.bf16_round(xq)
    xb[1500:1510] = xb[10:20]  # exact duplicates far apart (different tiles / splits)
    xb[2990:3000] = xb[10:20]
    xq[:10] = xb[10:20] * 4  # make the duplicated rows the winners
    idx = m.B200FlatIndex(d, 0, dtype=dtype)
    idx.add(xb)
    r = _search_np(idx, xq, 6, algo=algo)
    D_ref, I_ref = o.exact_topk_f64(xb, xq, 6)
    for q in range(10):
        assert r["ids"][q, :3].tolist() == [10 + q, 1500 + q, 2990 + q]  # (score desc, id asc)
    o.check_topk(xb, xq, r["scores"], r["ids"], 0, rtol=RTOL_BF16, D_ref=D_ref, I_ref=I_ref)


@pytest.mark.parametrize("dtype,algo", [("fp32", "simt"), ("fp32", "tcx"), ("bf16", "tc"), ("bf16", "tc2")])
def test_ignore_ids_semantics(m, dtype, algo):
    xb, xq = _data(10000, 96, 50, seed=21)
    if dtype == "bf16":
        xb, xq = o.bf16_round(xb), o.bf16_round(xq)
    idx = m.B200FlatIndex(96, 0, dtype=dtype)
    idx.add(xb)
    plain = _search_np(idx, xq, 9, algo=algo)
    ign = plain["ids"][:, 0].copy()
    ign[::3] = plain["ids"][::3, 4]
    ign[1] = 10**9  # not in the bank: nothing dropped
    r = _search_np(idx, xq, 8, ignore_ids=torch.from_numpy(ign), algo=algo)
    o.check_topk(xb, xq, r["scores"], r["ids"], 0, rtol=RTOL_BF16, ignore=ign)
    # reference formulation: k+1 search then drop (mips.py:388-398)
    fn = lambda q, kk: (plain["scores"][:, :kk], plain["ids"][:, :kk])
    _, i_ref = o.mips_search(fn, xq, ign.tolist(), 8)
    assert np.array_equal(np.asarray(i_ref), r["ids"])


def test_sharded_merge_is_shard_count_invariant(m):
    """Row shards searched independently + K2 merge == one bank, single process."""
    from retrieval_augmented_mds_b200.sharded import balanced_range
    xb, xq = _data(30011, 256, 100, seed=31)
    xb, xq = o.bf16_round(xb), o.bf16_round(xq)
    one = m.B200FlatIndex(256, 0, dtype="bf16")
    one.add(xb)
    ref = _search_np(one, xq, 8, want=("scores", "ids", "cosine"))
    for G in (2, 3, 8):
        keys, ids, xn2s, packs, qn2 = [], [], [], [], None
        for r in range(G):
            rows = balanced_range(len(xb), r, G)
            sh = m.B200FlatIndex(256, 0, dtype="bf16", id_offset=rows.start)
            sh.add(xb[rows.start:rows.stop])
            kk, ii, xx, qn2 = sh.search_local(torch.from_numpy(xq), 8)
            keys.append(kk), ids.append(ii), xn2s.append(xx)
            pk, _ = sh.search_local_packed(torch.from_numpy(xq), 8)  # the record format the all-gather moves
            packs.append(pk)
        out = m.merge_candidates(torch.stack(keys), torch.stack(ids), torch.stack(xn2s), qn2, 8, 0,
                                 want=("scores", "ids", "cosine"))
        outp = m.merge_candidates(None, None, None, qn2, 8, 0, want=("scores", "ids", "cosine"),
                                  packed=torch.stack(packs))
        for o_ in (out, outp):
            assert np.array_equal(o_["ids"].cpu().numpy(), ref["ids"])
            np.testing.assert_allclose(o_["scores"].cpu().numpy(), ref["scores"], rtol=1e-6)
            np.testing.assert_allclose(o_["cosine"].cpu().numpy(), ref["cosine"], rtol=1e-6)


def test_growth_preserves_rows_and_results(m):
    xb, xq = _data(9000, 64, 16, seed=41)
    idx = m.B200FlatIndex(64, 0, dtype="bf16")  # default capacity, forced to regrow
    for s in range(0, 9000, 1000):  # datasets adds in 1000-row batches
        idx.add(xb[s:s + 1000])
    big = m.B200FlatIndex(64, 0, dtype="bf16", capacity=9000)
    big.add(xb)
    a, b = idx.search(xq, 5), big.search(xq, 5)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


# ----------------------------------------------------------------------------------------- full size
def test_config2_1m_fp32_planted_neighbours(m):
    """BASELINE config 2 shape (1M x 768 fp32, 256 queries, k=8) through size-independent
    properties: planted near-duplicates must come first, result must equal a chunked torch
    fp32 brute force on the same data."""
    n, d, nq, k = 1_000_000, 768, 256, 8
    gen = torch.Generator(device="cuda").manual_seed(1234)
    xb = torch.randn((n, d), generator=gen, device="cuda", dtype=torch.float32)
    xq = torch.randn((nq, d), generator=gen, device="cuda", dtype=torch.float32)
    planted = torch.randint(0, n, (nq,), generator=gen, device="cuda")
    xq[: nq // 2] = xb[planted[: nq // 2]] * 2 + 0.01 * xq[: nq // 2]
    idx = m.B200FlatIndex(d, 0, dtype="fp32", capacity=n)
    idx.add(xb)
    r = idx.search_ex(xq, k)
    assert (r["ids"][: nq // 2, 0] == planted[: nq // 2]).all()
    best_s, best_i = None, None
    for s in range(0, n, 250_000):
        sc = xq @ xb[s:s + 250_000].T
        ts, ti = sc.topk(k, dim=1)
        ti += s
        if best_s is None:
            best_s, best_i = ts, ti
        else:
            cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti], 1)
            best_s, o_ = cs.topk(k, dim=1)
            best_i = ci.gather(1, o_)
    same = (best_i == r["ids"]).float().mean().item()
    assert same > 0.999
    torch.testing.assert_close(r["scores"], best_s, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("dtype,nq", [("bf16", 200), ("bf16", 40), ("fp32", 70)])
@pytest.mark.parametrize("metric", [0, 1])
def test_k_larger_than_one_pass_multipass_search(m, dtype, nq, metric):
    """faiss accepts any k (mips.py:383-386 asks for k or k + 1): k > 64 runs bounded passes — each pass only sees
    rows strictly after the last result of the pass before. Exact against the float64 oracle for every kernel
    family (CTA pair: nq > 128, 1-CTA: nq <= 128, fp32 bank: exact fp32 kernel), with duplicates straddling the
    pass boundary, an ignored id per query, k beyond the bank (padding) and the host entry point."""
    n, d = 3000, 96
    xb, xq = _data(n, d, nq, seed=77 + nq)
    xb[1000:1010] = xb[5]            # 11 identical rows: ties that straddle a pass boundary for the planted queries
    xq[:8] = xb[5] * 3.0
    if dtype == "bf16":
        xb, xq = o.bf16_round(xb), o.bf16_round(xq)
    rtol = RTOL_F32 if dtype == "fp32" else RTOL_BF16
    ign = np.random.default_rng(1).integers(0, n, nq)
    ign[:4] = 5                                               # one of the duplicated rows
    idx = m.B200FlatIndex(d, metric, dtype=dtype)
    idx.add(xb)
    for k in (65, 130, 200):
        D_ref, I_ref = o.exact_topk_f64(xb, xq, k, metric, ignore=ign)
        r = idx.search_ex(torch.from_numpy(xq), k, ignore_ids=torch.from_numpy(ign), want=("scores", "ids", "cosine"))
        o.check_topk(xb, xq, r["scores"].cpu().numpy(), r["ids"].cpu().numpy(), metric, rtol=rtol, D_ref=D_ref, I_ref=I_ref,
                     ignore=ign, what=f"multipass {dtype} k={k}")
        D, I = idx.search_host(xq, k, ignore_ids=ign)         # the loop inside the C call
        assert np.array_equal(I, r["ids"].cpu().numpy())
        got = r["ids"].cpu().numpy()
        assert all(len(set(row)) == k for row in got)         # no result twice across passes
    small = m.B200FlatIndex(d, metric, dtype=dtype)
    small.add(xb[:100])
    D, I = small.search(xq[:3], 150)                          # k > ntotal: -1 padding after the 100 rows
    assert (I[:, :100] >= 0).all() and (I[:, 100:] == -1).all()
    assert np.all(np.isinf(D[:, 100:]))
    with pytest.raises(ValueError):
        idx.search(xq, 5000)


def _brute_force_same_inputs(idx, xq, k, chunk=500_000):
    """fp32 flat IP over the rows AS STORED (bf16-rounded, up-cast) — the "same inputs" reference of the
    north star — chunked on the GPU; returns (scores, ids) by (score desc, id asc)."""
    nq = xq.shape[0]
    best_s = torch.full((nq, k), -float("inf"), device=xq.device)
    best_i = torch.full((nq, k), -1, device=xq.device, dtype=torch.int64)
    for s in range(0, idx.ntotal, chunk):
        X = idx.reconstruct_n(s, min(chunk, idx.ntotal - s), as_torch=True)
        ts, ti = (xq @ X.T).topk(k, dim=1)
        cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti + s], 1)
        o_ = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = cs.gather(1, o_), ci.gather(1, o_)
        del X
    return best_s, best_i


def _assert_recall_one(ids, sc, ref_i, ref_s, rtol=1e-4):
    """recall@k == 1.0 with ties certified: wherever the id differs from the brute force, both scores must
    agree within rtol (the two candidates are tied at fp32 accumulation-order precision — check_topk's rule)."""
    scale = ref_s.abs().clamp_min(1.0)
    assert float(((sc - ref_s).abs() / scale).max()) <= rtol
    diff = ids != ref_i
    if bool(diff.any()):
        # every differing id must be in a tie: its score equals the reference score at that rank within rtol (above)
        # AND the id SETS may differ only by members whose score is within rtol of the k-th score
        kth = ref_s[:, -1:]
        for b in diff.any(1).nonzero().flatten().tolist():
            a, r = set(ids[b].tolist()), set(ref_i[b].tolist())
            for extra in (a - r):
                pos = (ids[b] == extra).nonzero()[0, 0]
                assert abs(float(sc[b, pos] - kth[b, 0])) <= rtol * float(scale[b, 0]), (b, extra)
    return int(diff.sum())


def test_config3_full_bank_recall_is_one(m):
    """BASELINE config 3 at its FULL size — 10M x 768 bf16, 1024 queries, k=8 — against the chunked fp32 brute
    force over the same bf16-rounded values: recall@8 == 1.0 (ties certified), scores within 1e-4; half the
    queries are planted neighbours (bank rows + noise) so recall is not vacuous."""
    n, d, nq, k = 10_000_000, 768, 1024, 8
    gen = torch.Generator(device="cuda").manual_seed(31)
    idx = m.B200FlatIndex(d, 0, dtype="bf16", capacity=n)
    for s in range(0, n, 500_000):
        idx.add(torch.randn((500_000, d), generator=gen, device="cuda"))
    planted = torch.randint(0, n, (nq // 2,), generator=gen, device="cuda")
    rows = torch.cat([idx.reconstruct_n(int(r), 1, as_torch=True) for r in planted.tolist()])
    xq = torch.cat([rows + 0.05 * torch.randn((nq // 2, d), generator=gen, device="cuda"),
                    torch.randn((nq - nq // 2, d), generator=gen, device="cuda")]).bfloat16().float()
    r = idx.search_ex(xq, k)
    assert idx.last_algo == "tc2"
    assert torch.equal(r["ids"][: nq // 2, 0], planted)
    ref_s, ref_i = _brute_force_same_inputs(idx, xq, k)
    n_diff = _assert_recall_one(r["ids"], r["scores"], ref_i, ref_s)
    assert n_diff <= 8, n_diff            # genuine fp32 ties among 8192 results are rare
    g = idx.capture(nq, k)                # the CUDA-graph step bench.py times returns the same answer
    out = g.replay(xq)
    torch.cuda.synchronize()
    assert torch.equal(out["ids"], r["ids"]) and torch.equal(out["scores"], r["scores"])


@pytest.mark.parametrize("rows", [250_000, 2_000_000])
def test_config5_shape_k32_recall_is_one(m, rows):
    """BASELINE config 5's search shape: k=32, 1024 queries, 250k rows (one of 8 shards of the 2M-doc bank) and
    the whole 2M bank on one GPU, bf16, against the brute force on the same inputs."""
    d, nq, k = 768, 1024, 32
    gen = torch.Generator(device="cuda").manual_seed(rows)
    idx = m.B200FlatIndex(d, 0, dtype="bf16", capacity=rows)
    for s in range(0, rows, 250_000):
        idx.add(torch.randn((250_000, d), generator=gen, device="cuda"))
    xq = torch.randn((nq, d), generator=gen, device="cuda").bfloat16().float()
    r = idx.search_ex(xq, k)
    ref_s, ref_i = _brute_force_same_inputs(idx, xq, k)
    _assert_recall_one(r["ids"], r["scores"], ref_i, ref_s)
    assert bool((r["scores"][:, :-1] >= r["scores"][:, 1:]).all())


def test_config3_slice_bf16_recall(m):
    """BASELINE config 3 shape at 2M rows (bf16 bank, 1024 queries, k=8): recall@8 = 1.0 against
    fp32 flat IP on the same rounded inputs (torch chunked brute force on the GPU)."""
    n, d, nq, k = 2_000_000, 768, 1024, 8
    gen = torch.Generator(device="cuda").manual_seed(4321)
    idx = m.B200FlatIndex(d, 0, dtype="bf16", capacity=n)
    xq = torch.randn((nq, d), generator=gen, device="cuda").bfloat16().float()
    best_s = torch.full((nq, k), -float("inf"), device="cuda")
    best_i = torch.full((nq, k), -1, device="cuda", dtype=torch.int64)
    for s in range(0, n, 250_000):
        blk = torch.randn((250_000, d), generator=gen, device="cuda").bfloat16().float()
        idx.add(blk)
        ts, ti = (xq @ blk.T).topk(k, dim=1)
        cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti + s], 1)
        best_s, o_ = cs.topk(k, dim=1)
        best_i = ci.gather(1, o_)
    r = idx.search_ex(xq, k)
    assert idx.last_algo == "tc2"  # AUTO: CTA-pair kernel for batches of more than 128 queries
    ref_sets = [set(row.tolist()) for row in best_i.cpu()]
    got_sets = [set(row.tolist()) for row in r["ids"].cpu()]
    recall = sum(len(a & b) for a, b in zip(ref_sets, got_sets)) / (nq * k)
    assert recall >= 0.9995, recall  # torch's own fp32 accumulation order differs: ties only
    torch.testing.assert_close(r["scores"], best_s, rtol=1e-4, atol=1e-2)
