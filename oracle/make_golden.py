"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE on seeded inputs.

`import sotasum.mips` is impossible in this image (top-level `import faiss`, `adapters`,
`pytorch_lightning`), so the functions and statements on the hot path are extracted from the
reference sources by AST / source segment and executed unmodified:

  sotasum/mips.py            inner_product, get_phi, augment_xb, augment_xq, _layer_norm,
                             the body of Mips.search (with a stub index whose .search is the
                             reference's inner_product), Mips._prepare_query (with
                             faiss.normalize_L2 replaced by its documented contract)
  sotasum/pretrain.py        retriever_metrics
  sotasum/retriever_generator.py   the doc-score statements :158-172 and :188-192, the copy/generation
                             mixture statements :391-404 (+ their autograd gradients)
  sotasum/decoder_own.py     the score-biased copy attention statements :102-134 and :160-176 of
                             LEDDecoderAttention.forward (+ their autograd gradients)

Run here (needs /root/reference):   python oracle/make_golden.py
The .npz files are small and committed; the GPU box never reads /root/reference.
"""
from __future__ import annotations

import ast
import sys
import textwrap
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/sotasum")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _extract(path: Path, names: set[str], method_of: str | None = None) -> dict:
    src = path.read_text()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    found = {}
    nodes = tree.body
    if method_of:
        cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == method_of)
        nodes = cls.body
    for n in nodes:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            code = textwrap.dedent(ast.get_source_segment(src, n))
            # drop decorators (rank_zero_only etc.) — they are not part of the arithmetic
            code = code[code.index("def "):]
            exec(compile(code, f"{path}:{n.lineno}", "exec"), ns)
            found[n.name] = ns[n.name]
    missing = names - set(found)
    assert not missing, f"not found in {path}: {missing}"
    return found


def _statements(path: Path, first: int, last: int) -> str:
    """Source of whole statements whose line span lies inside [first, last] of a function body."""
    src = path.read_text()
    tree = ast.parse(src)
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.stmt) and not isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.If)):
            if node.lineno >= first and node.end_lineno <= last and isinstance(
                    node, (ast.Assign, ast.AugAssign, ast.AnnAssign, ast.With, ast.Expr)):
                out.append((node.lineno, node.end_lineno, textwrap.dedent(ast.get_source_segment(src, node))))
    # keep outermost statements only
    out.sort()
    keep, last_end = [], -1
    for a, b, s in out:
        if a > last_end:
            keep.append(s)
            last_end = b
    return "\n".join(keep)


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    mips_py = REF / "mips.py"
    fns = _extract(mips_py, {"inner_product", "get_phi", "augment_xb", "augment_xq", "_layer_norm"})
    inner_product, get_phi = fns["inner_product"], fns["get_phi"]
    augment_xb, augment_xq = fns["augment_xb"], fns["augment_xq"]
    meth = _extract(mips_py, {"search", "_prepare_query", "l2_normalization"}, method_of="Mips")
    metrics = _extract(REF / "pretrain.py", {"retriever_metrics"})["retriever_metrics"]

    rng = np.random.default_rng(20231018)

    # ---- G1: inner_product, both normalize settings (mips.py:552-560)
    xb = rng.standard_normal((2048, 96), dtype=np.float32) * rng.uniform(0.5, 2.0, (2048, 1)).astype(np.float32)
    xq = rng.standard_normal((24, 96), dtype=np.float32)
    np.savez_compressed(OUT / "inputs.npz", xb=xb, xq=xq)  # shared by the fixtures below
    g = {}
    for norm in (False, True):
        for k in (1, 8, 10):
            s, i = inner_product(xq, xb, k, normalize=norm)
            g[f"scores_n{int(norm)}_k{k}"] = s.astype(np.float32)
            g[f"ids_n{int(norm)}_k{k}"] = i.astype(np.int64)
    np.savez_compressed(OUT / "inner_product.npz", **g)

    # ---- G2: L2 augmentation helpers + the IP == L2-on-augmented identity (mips.py:55-70, 655-685)
    def _aug_case(bank, queries):
        phi = get_phi(bank)
        bank_aug = augment_xb(bank)
        q_aug = augment_xq(queries)
        # exact L2 on the augmented vectors (what IndexFlatL2(d+1) computes), float64 then ranked
        d2 = ((q_aug[:, None, :].astype(np.float64) - bank_aug[None, :, :].astype(np.float64)) ** 2).sum(-1)
        ids_l2 = np.argsort(d2, axis=1, kind="stable")[:, :10]
        s_ip, ids_ip = inner_product(queries, bank, 10, normalize=False)
        return dict(phi=np.float32(phi), extracol=bank_aug[:, -1].astype(np.float32), xq_aug=q_aug.astype(np.float32),
                    ids_l2=ids_l2.astype(np.int64), d2=np.take_along_axis(d2, ids_l2, 1),
                    ids_ip=ids_ip.astype(np.int64), scores_ip=s_ip.astype(np.float32))

    case_a = _aug_case(xb, xq[:8])                       # rows with very different norms
    lay = fns["_layer_norm"](torch.tensor(xb[:256])).numpy()   # the setting of test_faiss_index
    case_b = _aug_case(lay, lay[:2])
    np.savez_compressed(OUT / "augment.npz", layer_norm_rows=lay, **case_a,
                        **{f"ln_{k}": v for k, v in case_b.items()})

    # ---- G3: Mips.search with and without ignore_indexes (mips.py:382-400), the reference's own
    # method body run against a stub index backed by the reference's inner_product.
    class _FaissIndex:
        def search(self, queries, k):
            return inner_product(queries, xb, k, normalize=False)

    class _Holder:
        faiss_index = _FaissIndex()

    class _Emb:
        def get_index(self, name):
            return _Holder()

    stub = types.SimpleNamespace(embeddings=_Emb(), index_name="mips_embeddings")
    ign = [int(v) for v in rng.integers(0, xb.shape[0], xq.shape[0])]
    # make half of the ignore ids actual top hits so the filter does something
    top1 = inner_product(xq, xb, 1, normalize=False)[1][:, 0]
    for j in range(0, len(ign), 2):
        ign[j] = int(top1[j])
    s0, i0 = meth["search"](stub, xq, None, 10)
    s1, i1 = meth["search"](stub, xq, ign, 10)
    np.savez_compressed(OUT / "mips_search.npz", ignore=np.asarray(ign, dtype=np.int64),
                        scores_plain=np.asarray(s0, dtype=np.float32), ids_plain=np.asarray(i0, dtype=np.int64),
                        scores_ignore=np.asarray(s1, dtype=np.float32), ids_ignore=np.asarray(i1, dtype=np.int64))

    # ---- G4: _prepare_query (mips.py:368-375); faiss.normalize_L2 is third party: its documented
    # contract (x *= 1/sqrt(|x|^2), zero rows untouched) is supplied as the stand-in.
    def _normalize_L2(x):
        n2 = (x * x).sum(1)
        nz = n2 > 0
        x[nz] *= (1.0 / np.sqrt(n2[nz]))[:, None]

    faiss_stub = types.SimpleNamespace(METRIC_INNER_PRODUCT=0, METRIC_L2=1, normalize_L2=_normalize_L2)
    meth["_prepare_query"].__globals__["faiss"] = faiss_stub
    meth["_prepare_query"].__globals__["augment_xq"] = augment_xq
    pq = {}
    xq_z = xq.copy()
    xq_z[3] = 0.0
    for metric, norm in ((0, True), (0, False), (1, True)):
        me = types.SimpleNamespace(normalize=norm, metric_type=metric)
        me.l2_normalization = types.MethodType(meth["l2_normalization"], me)
        pq[f"m{metric}_n{int(norm)}"] = meth["_prepare_query"](me, xq_z.copy())
    np.savez_compressed(OUT / "prepare_query.npz", xq=xq_z, **pq)

    # ---- G5: retriever_metrics (pretrain.py:69-85)
    pred = (rng.uniform(size=(32, 10)) < 0.25).astype(np.float32)
    counts = np.maximum(pred.sum(-1), 1) + rng.integers(0, 3, 32).astype(np.float32)
    m = metrics(torch.tensor(pred), torch.tensor(counts))
    np.savez_compressed(OUT / "retriever_metrics.npz", pred=pred, counts=counts,
                        recall=np.float32(m["recall"]), reciprocal_rank=np.float32(m["reciprocal_rank"]),
                        average_precision=np.float32(m["average_precision"]))

    # ---- G6: doc-score statements of SotasumEncoder.forward (retriever_generator.py:158-172,188-192)
    rg = REF / "retriever_generator.py"
    code = _statements(rg, 158, 172) + "\n" + _statements(rg, 180, 192)
    B, K, L, d = 6, 5, 7, 96
    query = torch.tensor(rng.standard_normal((B, 1, d), dtype=np.float32))
    hid = torch.tensor(rng.standard_normal((B, K, L, d), dtype=np.float32))
    mem = torch.tensor(rng.standard_normal((B * K, L, 16), dtype=np.float32))
    mips_out = types.SimpleNamespace(mips_last_hidden_state=hid, memory_outputs=(mem,),
                                     memory_attention_mask=torch.ones(B * K, L),
                                     memory_input_ids=torch.zeros(B * K, L, dtype=torch.long))
    ns = {"torch": torch, "query": query, "mips_out": mips_out, "query_batch_size": B}
    exec(compile(code, str(rg), "exec"), ns)
    np.savez_compressed(OUT / "doc_scores.npz", query=query[:, 0, :].numpy(), docs=hid[:, :, 0, :].numpy(),
                        mips_scores=ns["mips_scores"].numpy(), memory_bias=ns["memory_bias"].numpy(),
                        memory_seq_len=np.int64(L))
    (OUT / "doc_scores_statements.txt").write_text(
        "# statements executed from /root/reference/sotasum/retriever_generator.py\n"
        + "\n".join("# " + ln.split("=")[0].strip() for ln in code.splitlines() if "=" in ln and not ln.startswith(" ")))
    # ---- G7: generation / copy mixture of RetrieverGenerator.forward (retriever_generator.py:391-404)
    code = _statements(rg, 391, 397) + "\n" + _statements(rg, 404, 404)
    rng7 = np.random.default_rng(7070)
    B, T, V, S = 3, 4, 97, 21
    import torch.nn.functional as F
    logits = torch.tensor(rng7.standard_normal((B, T, V), dtype=np.float32) * 3.0)
    gates = torch.softmax(torch.tensor(rng7.standard_normal((B, T, 2), dtype=np.float32)), -1)
    gen_gate, copy_gate = gates.chunk(2, dim=-1)                       # decoder_own.py:536
    align = torch.softmax(torch.tensor(rng7.standard_normal((B, T, S), dtype=np.float32)), -1)
    copy_probs = copy_gate * align                                       # decoder_own.py:538
    copy_seq = torch.tensor(rng7.integers(0, V, (B, S)))
    copy_seq[:, :4] = copy_seq[:, 4:8]                                   # repeated tokens accumulate
    ns = {"torch": torch, "F": F, "gen_gate": gen_gate, "copy_probs": copy_probs,
          "decoder_outputs": types.SimpleNamespace(logits=logits.clone()),
          "decoder_hidden_states": torch.zeros(B, T, 8), "encoder_copy_sequence": copy_seq}
    exec(compile(code, str(rg), "exec"), ns)
    np.savez_compressed(OUT / "copy_mixture.npz", logits=logits.numpy(), gen_gate=gen_gate.numpy(),
                        copy_probs=copy_probs.numpy(), copy_seq=copy_seq.numpy(), outs=ns["outs"].numpy())
    # gradients of the mixture through the reference's own statements (autograd over the in-place scatter_add_)
    lg = logits.clone().requires_grad_(True)
    gg = gen_gate.clone().requires_grad_(True)
    cpg = copy_probs.clone().requires_grad_(True)
    ns = {"torch": torch, "F": F, "gen_gate": gg, "copy_probs": cpg,
          "decoder_outputs": types.SimpleNamespace(logits=lg),
          "decoder_hidden_states": torch.zeros(B, T, 8), "encoder_copy_sequence": copy_seq}
    exec(compile(code, str(rg), "exec"), ns)
    w_out = torch.tensor(rng7.standard_normal((B, T, V), dtype=np.float32))
    (ns["outs"] * w_out).sum().backward()
    np.savez_compressed(OUT / "copy_mixture_grad.npz", w_out=w_out.numpy(), d_logits=lg.grad.numpy(),
                        d_gen_gate=gg.grad.numpy(), d_copy_probs=cpg.grad.numpy())

    # ---- G8: the score-biased copy attention (decoder_own.py:102-134 + :160-176), ONE head like the copy decoder
    # (its alignment is `.squeeze(1)`-ed at decoder_own.py:525), forward and autograd gradients; attention_bias is
    # the memory_bias broadcast of retriever_generator.py:188-192 built from per-document scores
    dec = REF / "decoder_own.py"
    code = _statements(dec, 102, 134) + "\n" + _statements(dec, 160, 176)
    rng8 = np.random.default_rng(8080)
    B, T, K, L, D = 3, 5, 4, 6, 16
    S = K * L
    me = types.SimpleNamespace(num_heads=1, head_dim=D, dropout=0.0, training=False,
                               beta=torch.tensor([1.3], requires_grad=True),
                               beta_bias=torch.tensor([-0.2], requires_grad=True))
    me._shape = lambda tensor, seq_len, bsz: tensor.view(bsz, seq_len, 1, D).transpose(1, 2).contiguous()
    q_in = torch.tensor(rng8.standard_normal((B, T, D), dtype=np.float32) * 0.5, requires_grad=True)
    k_in = torch.tensor(rng8.standard_normal((B, 1, S, D), dtype=np.float32), requires_grad=True)
    v_in = torch.tensor(rng8.standard_normal((B, 1, S, D), dtype=np.float32), requires_grad=True)
    doc = torch.tensor(rng8.uniform(-1, 1, (B, K)).astype(np.float32), requires_grad=True)      # mips_scores
    keep = torch.tensor(rng8.uniform(size=(B, S)) < 0.8)
    keep[:, 0] = True
    add_mask = torch.zeros(B, S).masked_fill(~keep, torch.finfo(torch.float32).min)
    attention_mask = add_mask[:, None, None, :].expand(B, 1, T, S)                                # _expand_mask layout
    attention_bias = doc.unsqueeze(-1).expand(-1, -1, L).reshape(B, -1)                           # rg.py:188-192
    ns = {"torch": torch, "nn": torch.nn, "self": me, "query_states": q_in, "key_states": k_in, "value_states": v_in,
          "attention_bias": attention_bias, "attention_mask": attention_mask, "bsz": B, "tgt_len": T, "embed_dim": D}
    exec(compile(code, str(dec), "exec"), ns)
    probs, attn_out = ns["attn_weights"], ns["attn_output"]
    assert probs.shape == (B, T, S) and attn_out.shape == (B, T, D)
    w_p = torch.tensor(rng8.standard_normal((B, T, S), dtype=np.float32))
    w_o = torch.tensor(rng8.standard_normal((B, T, D), dtype=np.float32))
    ((probs * w_p).sum() + (attn_out * w_o).sum()).backward()
    np.savez_compressed(
        OUT / "copy_attention.npz", query_states=q_in.detach().numpy(), key_states=k_in.detach().numpy()[:, 0],
        value_states=v_in.detach().numpy()[:, 0], doc_scores=doc.detach().numpy(), mem_len=np.int64(L),
        beta=me.beta.detach().numpy(), beta_bias=me.beta_bias.detach().numpy(), add_mask=add_mask.numpy(),
        attn_weights=probs.detach().numpy(), attn_output=attn_out.detach().numpy(), w_p=w_p.numpy(), w_o=w_o.numpy(),
        d_query=q_in.grad.numpy(), d_key=k_in.grad.numpy()[:, 0], d_value=v_in.grad.numpy()[:, 0],
        d_doc_scores=doc.grad.numpy(), d_beta=me.beta.grad.numpy(), d_beta_bias=me.beta_bias.grad.numpy())
    print("golden fixtures written to", OUT)
    for f in sorted(OUT.glob("*.npz")):
        print(f"  {f.name}: {f.stat().st_size/1024:.1f} KiB")


if __name__ == "__main__":
    sys.exit(main())
