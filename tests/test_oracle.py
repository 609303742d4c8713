"""The oracle (oracle/mips_oracle.py) against outputs of the reference's own code
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import mips_oracle as o


def test_inner_product_matches_reference(golden):
    g, inp = golden["inner_product"], golden["inputs"]
    xb, xq = inp["xb"], inp["xq"]
    for norm in (0, 1):
        for k in (1, 8, 10):
            s, i = o.inner_product(xq, xb, k, normalize=bool(norm))
            assert np.array_equal(i, g[f"ids_n{norm}_k{k}"])
            np.testing.assert_allclose(s, g[f"scores_n{norm}_k{k}"], rtol=1e-6, atol=1e-6)


def test_flat_search_ip_equals_reference_inner_product(golden):
    g, inp = golden["inner_product"], golden["inputs"]
    D, I = o.flat_search(inp["xb"], inp["xq"], 10, o.METRIC_INNER_PRODUCT, chunk=500)
    assert np.array_equal(I, g["ids_n0_k10"])
    np.testing.assert_allclose(D, g["scores_n0_k10"], rtol=1e-5, atol=1e-5)


def _layer_norm(x, eps=1e-12):  # mips.py:45-49 (only used by the reference's smoke test)
    u = x.mean(-1, keepdims=True)
    s = ((x - u) ** 2).mean(-1, keepdims=True)
    return (x - u) / np.sqrt(s + eps)


def test_augmentation_and_ip_l2_identity(golden):
    """mips.py:55-70 and the invariant of test_faiss_index (mips.py:655-685)."""
    g, inp = golden["augment"], golden["inputs"]
    lay = g["layer_norm_rows"]
    np.testing.assert_allclose(_layer_norm(inp["xb"][:256]).astype(np.float32), lay, rtol=1e-4, atol=1e-5)
    for prefix, bank, queries in (("", inp["xb"], inp["xq"][:8]), ("ln_", lay, lay[:2])):
        assert np.isclose(o.get_phi(bank), g[prefix + "phi"], rtol=1e-6)
        aug = o.augment_xb(bank)
        np.testing.assert_allclose(aug[:, -1], g[prefix + "extracol"], rtol=1e-6, atol=1e-6)
        np.testing.assert_array_equal(o.augment_xq(queries), g[prefix + "xq_aug"])
        # L2 on augmented == IP on originals (ids), via the flat-search contract
        D2, I2 = o.flat_search(aug, o.augment_xq(queries), 10, o.METRIC_L2)
        Dip, Iip = o.flat_search(bank, queries, 10, o.METRIC_INNER_PRODUCT)
        assert np.array_equal(Iip, g[prefix + "ids_ip"])
        if prefix == "":
            # layer-normed rows all have |x|^2 = d up to rounding: their extra column is pure
            # cancellation noise, so only the well-conditioned case pins ids/distances of L2
            assert np.array_equal(I2, g["ids_l2"])
            assert np.array_equal(I2, Iip)
            np.testing.assert_allclose(D2, g["d2"], rtol=1e-4, atol=1e-3)
        else:
            assert np.array_equal(g["ln_ids_l2"], g["ln_ids_ip"])  # the reference's own printout
        # |q~-x~|^2 = |q|^2 + phi - 2<q,x>  (what MIPS_OUT_AUGL2 computes without the column)
        qn = (queries ** 2).sum(1, keepdims=True)
        np.testing.assert_allclose(qn + g[prefix + "phi"] - 2 * Dip, g[prefix + "d2"], rtol=1e-4, atol=2e-3)


def test_mips_search_filter_matches_reference(golden):
    g, inp = golden["mips_search"], golden["inputs"]
    xb, xq = inp["xb"], inp["xq"]
    fn = lambda q, k: o.flat_search(xb, q, k, o.METRIC_INNER_PRODUCT)
    s0, i0 = o.mips_search(fn, xq, None, 10)
    assert isinstance(s0, np.ndarray) and np.array_equal(i0, g["ids_plain"])
    s1, i1 = o.mips_search(fn, xq, g["ignore"].tolist(), 10)
    assert isinstance(s1, list) and isinstance(i1, list)
    assert np.array_equal(np.asarray(i1), g["ids_ignore"])
    np.testing.assert_allclose(np.asarray(s1), g["scores_ignore"], rtol=1e-5, atol=1e-5)
    # equivalent formulation used by the kernels: mask the ignored id, then top-k
    D, I = o.exact_topk_f64(xb, xq, 10, ignore=g["ignore"])
    assert np.array_equal(I, g["ids_ignore"])


def test_prepare_query_matches_reference(golden):
    g = golden["prepare_query"]
    for metric, norm in ((0, True), (0, False), (1, True)):
        out = o.prepare_query(g["xq"].copy(), metric, norm)
        ref = g[f"m{metric}_n{int(norm)}"]
        assert out.shape == ref.shape and out.dtype == np.float32 and out.flags.c_contiguous
        np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-7)
    assert np.all(o.prepare_query(g["xq"].copy(), 0, True)[3] == 0)  # zero row untouched


def test_retriever_metrics_matches_reference(golden):
    g = golden["retriever_metrics"]
    m = o.retriever_metrics(g["pred"], g["counts"])
    for key in ("recall", "reciprocal_rank", "average_precision"):
        assert np.isclose(m[key], float(g[key]), rtol=1e-5), key


def test_doc_scores_match_reference(golden):
    g = golden["doc_scores"]
    s = o.doc_scores(g["query"], g["docs"])
    np.testing.assert_allclose(s, g["mips_scores"], rtol=1e-5, atol=1e-6)
    mb = o.memory_bias(g["mips_scores"], int(g["memory_seq_len"]))
    np.testing.assert_array_equal(mb, g["memory_bias"])
    p = o.doc_prob(s)
    np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-5)


def test_build_index_order_of_operations():
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((300, 32)).astype(np.float32) * 3
    bank, max_norm, phi = o.build_index(emb, o.METRIC_INNER_PRODUCT, True)
    assert phi is None and np.isclose(max_norm, np.linalg.norm(emb, axis=1).max())
    np.testing.assert_allclose(np.linalg.norm(bank, axis=1), 1.0, rtol=1e-5)
    bank, max_norm, phi = o.build_index(emb, o.METRIC_L2, True)
    assert bank.shape == (300, 33) and np.isclose(phi, (emb ** 2).sum(1).max(), rtol=1e-6)
    np.testing.assert_allclose((bank ** 2).sum(1), phi, rtol=1e-4)


def test_shard_range_partition():
    for n, g in ((10, 3), (100000, 8), (7, 8), (1000, 1)):
        rows = [list(o.shard_range(n, r, g)) for r in range(g)]
        flat = [x for r in rows for x in r if x < n]
        if n >= g:
            assert flat == list(range(n))


def test_bf16_round_is_rne():
    import torch
    x = torch.randn(4096, dtype=torch.float32) * 100
    assert np.array_equal(o.bf16_round(x.numpy()), x.bfloat16().float().numpy())


def test_check_topk_detects_errors():
    rng = np.random.default_rng(1)
    xb = rng.standard_normal((500, 16)).astype(np.float32)
    xq = rng.standard_normal((4, 16)).astype(np.float32)
    D, I = o.flat_search(xb, xq, 5)
    assert o.check_topk(xb, xq, D, I) == 0
    bad = I.copy()
    bad[0, 4] = int(np.setdiff1d(np.arange(500), I[0])[0])
    with pytest.raises(AssertionError):
        o.check_topk(xb, xq, D, bad)
    with pytest.raises(AssertionError):
        o.check_topk(xb, xq, D * 1.01, I)
    # an exact duplicate row is an acceptable swap (tie)
    xb2 = np.vstack([xb, xb[I[1, 0]][None]])
    D2, I2 = o.flat_search(xb2, xq, 5)
    swap = I2.copy()
    pos = np.where(I2[1] == 500)[0]
    assert len(pos) == 1
    a, b = np.where(I2[1] == I[1, 0])[0][0], pos[0]
    swap[1, a], swap[1, b] = I2[1, b], I2[1, a]
    assert o.check_topk(xb2, xq, D2, swap) == 2


def test_copy_mixture_matches_reference_statements(golden):
    """Golden: retriever_generator.py:391-404 executed on seeded tensors (oracle/make_golden.py G7)."""
    g = golden["copy_mixture"]
    out = o.copy_mixture(g["logits"], g["gen_gate"], g["copy_probs"], g["copy_seq"])
    np.testing.assert_allclose(out, g["outs"], rtol=2e-6, atol=2e-6)
    assert np.all(np.isfinite(out))


def test_chunked_cpu_search_equals_flat_search():
    """The torch sgemm + top-k + merge leg that bench.py times as the CPU baseline returns what the
    restated flat search returns (ids identical, (score desc, id asc) across chunk boundaries)."""
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((4099, 48), dtype=np.float32)
    xb[1000] = xb[3000]                      # an exact tie across two chunks
    xq = rng.standard_normal((21, 48), dtype=np.float32)
    xq[0] = xb[1000]
    D1, I1 = o.flat_search(xb, xq, 8)
    D2, I2 = o.flat_search_chunked(xb, xq, 8, chunk_rows=1500)
    assert np.array_equal(I1, I2)
    np.testing.assert_allclose(D1, D2, rtol=1e-6, atol=1e-5)
    D3, I3 = o.flat_search_chunked(xb[:5], xq, 8)
    assert (I3[:, 5:] == -1).all() and np.isinf(D3[:, 5:]).all()


def test_copy_attention_oracle_matches_reference_statements(golden):
    """decoder_own.py:102-134,160-176 executed on seeded tensors (forward + autograd) vs the restatement."""
    g = golden["copy_attention"]
    L, beta, bb = int(g["mem_len"]), float(g["beta"][0]), float(g["beta_bias"][0])
    out, p = o.copy_attention(g["query_states"], g["key_states"], g["value_states"], g["doc_scores"], L, beta, bb,
                              g["add_mask"])
    np.testing.assert_allclose(p, g["attn_weights"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(out, g["attn_output"], rtol=2e-5, atol=1e-6)
    assert (p[g["add_mask"][:, None, :].repeat(p.shape[1], 1) < 0] == 0).all()       # masked tokens get no mass
    gr = o.copy_attention_grad(g["query_states"], g["key_states"], g["value_states"], g["doc_scores"], L, beta, bb,
                               g["add_mask"], g["w_o"], g["w_p"])
    for name in ("d_query", "d_key", "d_value", "d_doc_scores"):
        np.testing.assert_allclose(gr[name], g[name], rtol=2e-4, atol=2e-5, err_msg=name)
    np.testing.assert_allclose(gr["d_beta"], g["d_beta"][0], rtol=2e-4, atol=2e-5)
    assert abs(gr["d_beta_bias"]) < 1e-9 and abs(float(g["d_beta_bias"][0])) < 1e-5     # a constant logit: no gradient


def test_copy_mixture_grad_oracle_matches_reference_autograd(golden):
    g, gg = golden["copy_mixture"], golden["copy_mixture_grad"]
    dz, dgate, dcopy = o.copy_mixture_grad(g["logits"], g["gen_gate"], g["copy_probs"], g["copy_seq"], gg["w_out"])
    np.testing.assert_allclose(dz, gg["d_logits"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(dgate, gg["d_gen_gate"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(dcopy, gg["d_copy_probs"], rtol=2e-4, atol=2e-3)
