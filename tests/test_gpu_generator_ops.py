"""GPU tests of the training-path operators at the end of the retrieval marginalisation: the
score-biased copy attention (sotasum/decoder_own.py:102-134,160-176) and the generation / copy mixture
(sotasum/retriever_generator.py:391-404), forward AND backward, against (a) golden tensors produced by
executing the reference's own statements under torch autograd (oracle/make_golden.py) and (b) the float64
oracle at the retriever-generator step's shapes (BASELINE config 4: batch 16, k=5, L=512, BART vocabulary)."""
import numpy as np
import pytest
import torch

import retrieval_augmented_mds_b200 as pkg
from oracle import mips_oracle as o

pytestmark = pytest.mark.gpu


def _t(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().requires_grad_(grad)


def test_copy_attention_golden_forward_and_gradients(cuda_device, golden):
    g = golden["copy_attention"]
    q, k, v, doc = (_t(g[n], True) for n in ("query_states", "key_states", "value_states", "doc_scores"))
    beta = torch.nn.Parameter(_t(g["beta"]))
    beta_bias = torch.nn.Parameter(_t(g["beta_bias"]))
    mask4 = _t(g["add_mask"])[:, None, None, :].expand(-1, 1, q.shape[1], -1)      # the reference's [B, 1, T, S]
    out, probs = pkg.copy_attention(q, k, v, doc, int(g["mem_len"]), beta, beta_bias, mask4)
    np.testing.assert_allclose(probs.detach().cpu().numpy(), g["attn_weights"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["attn_output"], rtol=1e-5, atol=1e-6)
    ((probs * _t(g["w_p"])).sum() + (out * _t(g["w_o"])).sum()).backward()
    for name, t in (("d_query", q), ("d_key", k), ("d_value", v), ("d_doc_scores", doc), ("d_beta", beta)):
        np.testing.assert_allclose(t.grad.cpu().numpy(), g[name], rtol=2e-4, atol=2e-5, err_msg=name)
    assert abs(float(beta_bias.grad)) < 1e-5 and abs(float(g["d_beta_bias"][0])) < 1e-5
    # the reference's general per-token attention_bias (= memory_bias, already broadcast): mem_len = 1
    bias_tok = _t(np.repeat(g["doc_scores"], int(g["mem_len"]), axis=1), True)
    p2 = pkg.biased_softmax(torch.bmm(q.detach(), k.detach().transpose(1, 2)), bias_tok, float(g["beta"][0]),
                            float(g["beta_bias"][0]), _t(g["add_mask"]), 1)
    assert torch.equal(p2, probs.detach())
    (p2 * _t(g["w_p"])).sum().backward()
    assert bias_tok.grad.shape == bias_tok.shape


def test_copy_attention_step_shape_matches_oracle(cuda_device):
    """Retriever-generator step shape: B=16, k=5 documents of L=512 tokens (S=2560), D=1024 (BART-large width),
    ragged documents (masked padding), gradients flow to the document scores."""
    rng = np.random.default_rng(4)
    B, T, K, L, D = 16, 24, 5, 512, 1024
    S = K * L
    q = rng.standard_normal((B, T, D), dtype=np.float32) * D ** -0.5
    k = rng.standard_normal((B, S, D), dtype=np.float32)
    v = rng.standard_normal((B, S, D), dtype=np.float32)
    doc = rng.uniform(-1, 1, (B, K)).astype(np.float32)
    lens = rng.integers(1, L + 1, (B, K))
    keep = (np.arange(L)[None, None, :] < lens[:, :, None]).reshape(B, S)
    add_mask = np.where(keep, 0.0, np.finfo(np.float32).min).astype(np.float32)
    w_o = rng.standard_normal((B, T, D), dtype=np.float32)
    w_p = rng.standard_normal((B, T, S), dtype=np.float32)
    tq, tk, tv, td = _t(q, True), _t(k, True), _t(v, True), _t(doc, True)
    out, probs = pkg.copy_attention(tq, tk, tv, td, L, 1.7, 0.3, _t(add_mask))
    ref_out, ref_p = o.copy_attention(q, k, v, doc, L, 1.7, 0.3, add_mask)
    np.testing.assert_allclose(probs.detach().cpu().numpy(), ref_p, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref_out, rtol=2e-4, atol=2e-5)
    assert float(probs.detach()[~torch.from_numpy(keep).cuda()[:, None, :].expand(-1, T, -1)].abs().max()) == 0.0
    ((probs * _t(w_p)).sum() + (out * _t(w_o)).sum()).backward()
    gr = o.copy_attention_grad(q, k, v, doc, L, 1.7, 0.3, add_mask, w_o, w_p)
    for name, t in (("d_query", tq), ("d_key", tk), ("d_value", tv), ("d_doc_scores", td)):
        ref = gr[name]
        np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=2e-3, atol=2e-4 * max(1.0, float(np.abs(ref).max())),
                                   err_msg=name)


def test_copy_mixture_gradients_golden_and_bart_vocabulary(cuda_device, golden):
    g, gg = golden["copy_mixture"], golden["copy_mixture_grad"]
    logits, gate, cp = (_t(g[n], True) for n in ("logits", "gen_gate", "copy_probs"))
    out = pkg.copy_mixture(logits, gate, cp, _t(g["copy_seq"]))
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["outs"], rtol=1e-5, atol=1e-5)
    (out * _t(gg["w_out"])).sum().backward()
    np.testing.assert_allclose(logits.grad.cpu().numpy(), gg["d_logits"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(gate.grad.cpu().numpy(), gg["d_gen_gate"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(cp.grad.cpu().numpy(), gg["d_copy_probs"], rtol=2e-4, atol=2e-3)
    # BART vocabulary, k * L = 2560 memory positions, tokens repeated and one outside the vocabulary:
    # against torch autograd over the reference's own sequence of ops on the GPU, and the float64 oracle
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    B, T, V, S = 2, 3, 50265, 2560
    lg = (torch.randn((B, T, V), generator=gen, device=cuda_device) * 4).requires_grad_(True)
    gates = torch.softmax(torch.randn((B, T, 2), generator=gen, device=cuda_device), -1)
    gt = gates[..., :1].clone().requires_grad_(True)
    cpr = (gates[..., 1:] * torch.softmax(torch.randn((B, T, S), generator=gen, device=cuda_device), -1)).requires_grad_(True)
    seq = torch.randint(0, V, (B, S), generator=gen, device=cuda_device)
    seq[:, :100] = seq[:, 100:200]
    w = torch.randn((B, T, V), generator=gen, device=cuda_device)
    out = pkg.copy_mixture(lg, gt, cpr, seq)
    (out * w).sum().backward()
    mine = [t.grad.clone() for t in (lg, gt, cpr)]
    for t in (lg, gt, cpr):
        t.grad = None
    probs = gt * torch.softmax(lg, -1)
    probs = probs.scatter_add(-1, seq.reshape(B, 1, -1).expand(-1, T, -1), cpr)
    (torch.log(probs + 1e-7) * w).sum().backward()
    for a, t, name in zip(mine, (lg, gt, cpr), ("d_logits", "d_gen_gate", "d_copy_probs")):
        torch.testing.assert_close(a, t.grad, rtol=2e-3, atol=2e-4 * max(1.0, float(t.grad.abs().max())), msg=name)
    dz, dgate, dcopy = o.copy_mixture_grad(lg.detach().cpu().numpy(), gt.detach().cpu().numpy(), cpr.detach().cpu().numpy(),
                                           seq.cpu().numpy(), w.cpu().numpy())
    np.testing.assert_allclose(mine[0].cpu().numpy(), dz, rtol=2e-3, atol=2e-4 * float(np.abs(dz).max()))
    np.testing.assert_allclose(mine[1].cpu().numpy(), dgate, rtol=2e-3, atol=2e-4 * float(np.abs(dgate).max()))
    np.testing.assert_allclose(mine[2].cpu().numpy(), dcopy, rtol=2e-3, atol=2e-4 * float(np.abs(dcopy).max()))


def test_marginalisation_end_to_end_gradient_reaches_the_query(cuda_device):
    """search -> gathered rows -> cosine doc scores (with gradient w.r.t. the query) -> biased copy attention ->
    mixture -> loss: the retriever's query receives a gradient through the document scores, as in the
    reference's training step (retriever_generator.py:158-172 -> decoder_own.py:110-114 -> :391-404)."""
    rng = np.random.default_rng(9)
    n, d, B, K, L, T, V = 4000, 256, 4, 5, 16, 6, 1000
    bank = rng.standard_normal((n, d), dtype=np.float32)
    idx = pkg.B200FlatIndex(d, 0, dtype="fp32")
    idx.add(bank)
    query = _t(rng.standard_normal((B, d), dtype=np.float32), True)
    r = idx.search_ex(query.detach(), K)
    rows = idx.gather_rows(r["ids"])                                   # [B, K, d], frozen memory encoder
    cos = (query[:, None, :] * rows).sum(-1) / (query.detach().norm(dim=1, keepdim=True) * rows.norm(dim=2))
    hq = _t(rng.standard_normal((B, T, 64), dtype=np.float32) * 0.1)
    hk = _t(rng.standard_normal((B, K * L, 64), dtype=np.float32))
    out, align = pkg.copy_attention(hq, hk, hk, cos, L, 1.0, 0.0)
    gates = torch.softmax(_t(rng.standard_normal((B, T, 2), dtype=np.float32)), -1)
    seq = torch.randint(0, V, (B, K * L), device=cuda_device)
    logp = pkg.copy_mixture(_t(rng.standard_normal((B, T, V), dtype=np.float32)), gates[..., :1], gates[..., 1:] * align, seq)
    target = torch.randint(0, V, (B, T), device=cuda_device)
    loss = -logp.gather(-1, target[..., None]).mean()
    loss.backward()
    assert query.grad is not None and float(query.grad.abs().sum()) > 0.0 and bool(torch.isfinite(query.grad).all())
