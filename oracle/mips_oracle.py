"""CPU oracle for the MIPS hot path — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's exact-search arithmetic. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker or the reported CPU baseline — never as a product path.

Pinning status: the reference ships no tests, golden vectors or known-answer files for this
path. The oracle is therefore pinned against OUTPUTS OF THE REFERENCE'S OWN
CODE run in the build container: oracle/make_golden.py AST-extracts `inner_product`,
`get_phi`, `augment_xb`, `augment_xq` (sotasum/mips.py:55-70, 552-560), the body of
`Mips.search` (mips.py:382-400), `retriever_metrics` (pretrain.py:69-85) and the doc-score
statements of `SotasumEncoder.forward` (retriever_generator.py:158-172, 188-192) and the
generation / copy mixture statements (retriever_generator.py:391-404, with autograd gradients) and the
score-biased copy attention statements of LEDDecoderAttention.forward (decoder_own.py:102-134, 160-176,
with autograd gradients), executes
them on seeded inputs and commits the results under tests/golden/. tests/test_oracle.py checks
every function below against those fixtures. The faiss-cpu 1.7.4 kernels behind
`faiss_index.search` are a third-party wheel that is not vendored in /root/reference and not
installable here; their documented contract (IP: descending scores; L2: ascending squared
distances; int64 ids; -1 padding; normalize_L2 leaves zero rows untouched) is restated in
`flat_search` / `normalize_L2` and anchored on the reference's own invariant in
`test_faiss_index` (mips.py:655-685): ids of L2-on-augmented == ids of IP.
"""
from __future__ import annotations

import numpy as np

METRIC_INNER_PRODUCT = 0  # faiss.METRIC_INNER_PRODUCT (mips.py:306,369)
METRIC_L2 = 1  # faiss.METRIC_L2 (mips.py:316,371)


# ----------------------------------------------------------------------------------------------
# mips.py:55-70 — MIPS -> L2 reduction (Bachrach et al., theorem 5)
def get_phi(xb: np.ndarray):
    """mips.py:55-56: phi = max_i |x_i|^2."""
    return (xb**2).sum(1).max()


def augment_xb(xb: np.ndarray, phi=None) -> np.ndarray:
    """mips.py:59-65: append sqrt(phi - |x|^2) to every bank row."""
    norms = (xb**2).sum(1)
    if phi is None:
        phi = norms.max()
    extracol = np.sqrt(phi - norms)
    return np.hstack((xb, extracol.reshape(-1, 1)))


def augment_xq(xq: np.ndarray) -> np.ndarray:
    """mips.py:68-70: append a zero column to every query."""
    extracol = np.zeros(len(xq), dtype="float32")
    return np.hstack((xq, extracol.reshape(-1, 1)))


# ----------------------------------------------------------------------------------------------
def normalize_L2(x: np.ndarray) -> np.ndarray:
    """faiss.normalize_L2 as used at mips.py:521-525: x *= 1/sqrt(|x|^2) per row, rows with zero
    norm untouched. Returns a new float32 array (faiss works in place)."""
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    n2 = (x.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    nz = n2 > 0
    inv = np.ones_like(n2)
    inv[nz] = np.float32(1.0) / np.sqrt(n2[nz], dtype=np.float32)
    return (x * inv[:, None]).astype(np.float32)


def prepare_query(query: np.ndarray, metric_type: int, normalize: bool) -> np.ndarray:
    """Mips._prepare_query, mips.py:368-375."""
    if normalize and metric_type == METRIC_INNER_PRODUCT:
        query = normalize_L2(query)
    if metric_type == METRIC_L2:
        query = augment_xq(query)
    if not query.flags.c_contiguous:
        query = np.asarray(query, order="C")
    return query.astype(np.float32)


# ----------------------------------------------------------------------------------------------
def inner_product(x: np.ndarray, y: np.ndarray, k: int = 1, normalize: bool = True):
    """mips.py:552-560 (reached via Mips.np_search, :527-529): optional L2 normalisation of both
    sides, x @ y.T, full sort of the negated scores, first k. The reference's argsort is
    numpy's default introsort (tie order unspecified); the oracle uses a stable sort so ties
    resolve to the lower id, which is one of the orders the reference may produce."""
    assert len(x.shape) == len(y.shape) == 2
    if normalize:
        x = x / np.linalg.norm(x, axis=1, keepdims=True)
        y = y / np.linalg.norm(y, axis=1, keepdims=True)
    scores = x @ y.T
    indices = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    scores = np.take_along_axis(scores, indices, axis=1)
    return scores, indices


def flat_search(xb: np.ndarray, xq: np.ndarray, k: int, metric_type: int = METRIC_INNER_PRODUCT,
                chunk: int = 65536):
    """Contract of faiss.IndexFlatIP / IndexFlatL2 .search as called at mips.py:383-386:
    IP -> k largest <q,x>, descending; L2 -> k smallest |q-x|^2, ascending; ids int64; when
    k > ntotal the tail is id -1 with score -inf (IP) / +inf (L2). Chunked over the bank so the
    [nq, N] matrix is never fully materialised; ties resolve to the lower id."""
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    nq, n = xq.shape[0], xb.shape[0]
    best_s = np.full((nq, 0), 0, dtype=np.float32)
    best_i = np.full((nq, 0), 0, dtype=np.int64)
    qn = (xq**2).sum(1, dtype=np.float32)[:, None]
    for s in range(0, n, chunk):
        blk = xb[s:s + chunk]
        ip = xq @ blk.T
        if metric_type == METRIC_INNER_PRODUCT:
            key = ip
        else:
            bn = (blk**2).sum(1, dtype=np.float32)[None, :]
            key = -(np.maximum(qn + bn - 2.0 * ip, 0.0))
        ids = np.arange(s, s + blk.shape[0], dtype=np.int64)[None, :].repeat(nq, 0)
        cat_s = np.concatenate([best_s, key.astype(np.float32)], axis=1)
        cat_i = np.concatenate([best_i, ids], axis=1)
        # (key desc, id asc): ids are already ascending left to right, stable sort keeps that
        order = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]
        best_s = np.take_along_axis(cat_s, order, axis=1)
        best_i = np.take_along_axis(cat_i, order, axis=1)
    if best_s.shape[1] < k:
        pad = k - best_s.shape[1]
        best_s = np.concatenate([best_s, np.full((nq, pad), -np.inf, dtype=np.float32)], axis=1)
        best_i = np.concatenate([best_i, np.full((nq, pad), -1, dtype=np.int64)], axis=1)
    D = best_s if metric_type == METRIC_INNER_PRODUCT else -best_s
    return D.astype(np.float32), best_i


def merge_topk_pair(a, b, k: int):
    """Merge two (D, I) top-k results of disjoint row sets (IP: descending) by (score desc, id asc)."""
    cat_s = np.concatenate([a[0], b[0]], axis=1)
    cat_i = np.concatenate([a[1], b[1]], axis=1)
    order = np.lexsort((cat_i, -cat_s), axis=1)[:, :k]
    return np.take_along_axis(cat_s, order, axis=1), np.take_along_axis(cat_i, order, axis=1)


def flat_search_chunked(xb: np.ndarray, xq: np.ndarray, k: int, chunk_rows: int = 1_000_000):
    """The reference's CPU search path at full speed on all host threads, inner-product metric: what the flat
    faiss-cpu index behind `Mips.search` (mips.py:383-386) computes — sgemm of the query batch against a block
    of bank rows, the k best of each block, a running merge — spelled with the torch idiom the reference itself
    uses for brute-force scoring (`torch.topk(q @ x.T, k)`, retriever_lightning.py:304-305). Same result as
    `flat_search` (checked in tests/test_oracle.py); this is the leg bench.py times as the CPU baseline."""
    import torch
    tq = torch.from_numpy(np.ascontiguousarray(xq, dtype=np.float32))
    n = xb.shape[0]
    best = None
    for s in range(0, n, chunk_rows):
        blk = torch.from_numpy(xb[s:s + chunk_rows])
        v, i = (tq @ blk.T).topk(min(k, blk.shape[0]), dim=1)
        cur = (v.numpy(), i.numpy().astype(np.int64) + s)
        best = cur if best is None else merge_topk_pair(best, cur, k)
    D, I = best
    if D.shape[1] < k:
        pad = k - D.shape[1]
        D = np.concatenate([D, np.full((D.shape[0], pad), -np.inf, dtype=np.float32)], axis=1)
        I = np.concatenate([I, np.full((I.shape[0], pad), -1, dtype=np.int64)], axis=1)
    return D.astype(np.float32), I


def mips_search(search_fn, queries: np.ndarray, ignore_indexes=None, k: int = 10):
    """Mips.search, mips.py:382-400: fetch k (or k+1 with an ignore list), drop the hit whose id
    equals ignore_indexes[j], keep the first k. Returns numpy arrays without the filter and
    Python lists of lists with it, exactly like the reference."""
    scores, indices = search_fn(queries, k + 1 if ignore_indexes is not None else k)
    if ignore_indexes is not None:
        scores = [
            [s for i, s in enumerate(score) if ignore_indexes[j] != indices[j][i]][:k]
            for j, score in enumerate(scores)
        ]
        indices = [[i for i in index if ignore_indexes[j] != i][:k] for j, index in enumerate(indices)]
    return scores, indices


# ----------------------------------------------------------------------------------------------
def shard_range(n_rows: int, rank: int, num_rank: int) -> range:
    """Mips.encode_text2 partition rule, mips.py:226-230."""
    chunck_size = (n_rows // num_rank) + 1
    stop = (rank + 1) * chunck_size if rank + 1 < num_rank else n_rows
    return range(rank * chunck_size, stop)


def build_index(embeddings: np.ndarray, metric_type: int, normalize: bool):
    """Mips.build_index, mips.py:290-345, for string_factory "Flat": returns
    (bank_as_added_to_faiss, max_norm, phi). Order of operations as in the reference:
    max_norm over the raw rows (:298-304, _map_norm :347-349), then row normalisation when
    IP ∧ normalize (:306-314), then phi + augmentation when L2 (:316-331)."""
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    max_norm = np.linalg.norm(emb, axis=1, keepdims=True).max()
    phi = None
    if normalize and metric_type == METRIC_INNER_PRODUCT:
        emb = normalize_L2(emb)
    if metric_type == METRIC_L2:
        phi = (emb**2).sum(1).max()
        emb = augment_xb(emb, phi=phi)
    return emb.astype(np.float32), max_norm, phi


# ----------------------------------------------------------------------------------------------
def doc_scores(query: np.ndarray, docs: np.ndarray):
    """retriever_generator.py:158-172: mips_scores[b,j] = <q_b, d_bj> / (|q_b| |d_bj|).
    query [B, d], docs [B, k, d] -> [B, k]."""
    q = query.astype(np.float32)
    d = docs.astype(np.float32)
    s = np.einsum("bd,bkd->bk", q, d)
    qn = np.linalg.norm(q, axis=1, keepdims=True)
    dn = np.linalg.norm(d, axis=2)
    return (s / (qn * dn)).astype(np.float32)


def memory_bias(mips_scores: np.ndarray, memory_seq_len: int) -> np.ndarray:
    """retriever_generator.py:188-192: bias[b, j*L + t] = mips_scores[b, j]."""
    b, k = mips_scores.shape
    return np.broadcast_to(mips_scores[:, :, None], (b, k, memory_seq_len)).reshape(b, -1).copy()


def doc_prob(mips_scores: np.ndarray, beta: float = 1.0, beta_bias: float = 0.0) -> np.ndarray:
    """Per-document factor of the biased copy attention (decoder_own.py:110-114 adds
    beta*bias + beta_bias to the logits, :134 takes one softmax over all k*L memory tokens):
    with equal token logits the attention mass of doc j is softmax_j(beta*score_j + beta_bias)."""
    z = beta * mips_scores.astype(np.float64) + beta_bias
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def copy_mixture(logits: np.ndarray, gen_gate: np.ndarray, copy_probs: np.ndarray, copy_seq: np.ndarray,
                 eps: float = 1e-7) -> np.ndarray:
    """retriever_generator.py:391-404: log(gen_gate * softmax(logits) + scatter_add(copy_probs at
    copy_seq) + 1e-7). logits [B, T, V], gen_gate [B, T, 1], copy_probs [B, T, S] (already gated,
    decoder_own.py:538), copy_seq int [B, S] -> [B, T, V]. float32 like the reference."""
    x = logits.astype(np.float32)
    x = x - x.max(-1, keepdims=True)
    e = np.exp(x)
    probs = (gen_gate.astype(np.float32) * (e / e.sum(-1, keepdims=True))).astype(np.float32)
    B, T, _ = probs.shape
    for b in range(B):
        for t in range(T):
            np.add.at(probs[b, t], copy_seq[b], copy_probs[b, t].astype(np.float32))
    return np.log(probs + np.float32(eps)).astype(np.float32)


def copy_mixture_grad(logits, gen_gate, copy_probs, copy_seq, d_out, eps: float = 1e-7):
    """Gradients of `copy_mixture` (retriever_generator.py:391-404) for an upstream d_out [B, T, V], float64:
    with m = gen_gate * p + scatter(copy) + eps, dM = d_out / m: d_gate = sum dM p; d_logits = gate p (dM - sum dM p);
    d_copy[s] = dM[copy_seq[s]]. Returns (d_logits, d_gen_gate [B, T, 1], d_copy_probs)."""
    x = logits.astype(np.float64)
    x = x - x.max(-1, keepdims=True)
    p = np.exp(x)
    p /= p.sum(-1, keepdims=True)
    g = gen_gate.astype(np.float64).reshape(*logits.shape[:2], 1)
    m = g * p
    B, T, _ = m.shape
    for b in range(B):
        for t in range(T):
            np.add.at(m[b, t], copy_seq[b], copy_probs[b, t].astype(np.float64))
    dM = d_out.astype(np.float64) / (m + eps)
    dot = (dM * p).sum(-1, keepdims=True)
    d_copy = np.stack([dM[b][:, copy_seq[b]] for b in range(B)])
    return g * p * (dM - dot), dot, d_copy


def copy_attention(query_states, key_states, value_states, doc_scores, mem_len: int, beta: float = 1.0,
                   beta_bias: float = 0.0, add_mask=None):
    """The copy decoder's one-head cross attention, decoder_own.py:102-134 and :160-176: logits = q k^T (:108) +
    beta * attention_bias + beta_bias (:110-114, attention_bias[b, j * mem_len + t] = doc_scores[b, j] =
    memory_bias of retriever_generator.py:188-192) + additive mask (:123-132); ONE softmax over all memory tokens
    (:134); output = probs v (:164). q [B, T, D], k / v [B, S, D], doc_scores [B, K], add_mask [B, S].
    Returns (attn_output [B, T, D], attn_weights [B, T, S]) in float64."""
    q, k, v = (a.astype(np.float64) for a in (query_states, key_states, value_states))
    S = k.shape[1]
    z = q @ k.transpose(0, 2, 1)
    if doc_scores is not None:
        bias = np.repeat(doc_scores.astype(np.float64), mem_len, axis=1)[:, :S]
        z = z + (beta * bias + beta_bias)[:, None, :]
    if add_mask is not None:
        z = z + add_mask.astype(np.float64)[:, None, :]
    z = z - z.max(-1, keepdims=True)
    p = np.exp(z)
    p /= p.sum(-1, keepdims=True)
    return p @ v, p


def copy_attention_grad(query_states, key_states, value_states, doc_scores, mem_len: int, beta: float, beta_bias: float,
                        add_mask, d_out, d_probs):
    """Gradients of `copy_attention` for upstream d_out [B, T, D] (on attn_output) and d_probs [B, T, S] (on the
    alignment, which feeds copy_probs at decoder_own.py:538), float64. Returns a dict with d_query, d_key, d_value,
    d_doc_scores [B, K] (the retriever's signal: beta * sum over the document's tokens and all target positions of
    dS), d_beta, d_beta_bias."""
    q, k, v = (a.astype(np.float64) for a in (query_states, key_states, value_states))
    out, p = copy_attention(q, k, v, doc_scores, mem_len, beta, beta_bias, add_mask)
    dP = d_probs.astype(np.float64) + d_out.astype(np.float64) @ v.transpose(0, 2, 1)
    dS = p * (dP - (dP * p).sum(-1, keepdims=True))
    B, T, S = p.shape
    K = doc_scores.shape[1]
    pad = K * mem_len - S
    G = np.pad(dS.sum(1), ((0, 0), (0, pad))).reshape(B, K, mem_len).sum(-1)
    return {"d_query": dS @ k, "d_key": dS.transpose(0, 2, 1) @ q, "d_value": p.transpose(0, 2, 1) @ d_out.astype(np.float64),
            "d_doc_scores": beta * G, "d_beta": (G * doc_scores.astype(np.float64)).sum(), "d_beta_bias": G.sum()}


def retriever_metrics(pred: np.ndarray, counts: np.ndarray) -> dict:
    """pretrain.py:69-85 (copy at retriever_lightning.py:71-87), including the reference's
    reciprocal-rank quirk (1/argmax, inf -> 0: a rank-1 hit scores 0)."""
    pred = pred.astype(np.float32)
    counts = counts.astype(np.float32)
    recall = float((pred.sum(-1) / counts).mean())
    with np.errstate(divide="ignore"):
        rr = 1.0 / pred.argmax(-1).astype(np.float32)
    rr[np.isinf(rr)] = 0.0
    precision = (pred.cumsum(-1) / np.arange(1, pred.shape[-1] + 1, dtype=np.float32)) * pred
    ap = float((precision.sum(-1) / counts).mean())
    return {"recall": recall, "reciprocal_rank": float(rr.mean()), "average_precision": ap}


# ----------------------------------------------------------------------------------------------
# helpers for the parity tests
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round to nearest even) -> fp32, bit exact with torch / CUDA
    __float2bfloat16_rn for finite inputs."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def exact_topk_f64(xb: np.ndarray, xq: np.ndarray, k: int, metric_type: int = METRIC_INNER_PRODUCT,
                   ignore=None, chunk: int = 65536):
    """float64 ground truth with the deterministic (score desc / distance asc, id asc) rule and
    an optional per-query ignored id. Returns (D float64 [nq,k], I int64 [nq,k])."""
    xq64 = xq.astype(np.float64)
    nq, n = xq.shape[0], xb.shape[0]
    best_s = np.zeros((nq, 0))
    best_i = np.zeros((nq, 0), dtype=np.int64)
    qn = (xq64**2).sum(1)[:, None]
    for s in range(0, n, chunk):
        blk = xb[s:s + chunk].astype(np.float64)
        ip = xq64 @ blk.T
        key = ip if metric_type == METRIC_INNER_PRODUCT else -(qn + (blk**2).sum(1)[None, :] - 2 * ip)
        ids = np.arange(s, s + blk.shape[0], dtype=np.int64)[None, :].repeat(nq, 0)
        if ignore is not None:
            key = np.where(ids == np.asarray(ignore, dtype=np.int64)[:, None], -np.inf, key)
        cat_s = np.concatenate([best_s, key], axis=1)
        cat_i = np.concatenate([best_i, ids], axis=1)
        order = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]
        best_s = np.take_along_axis(cat_s, order, axis=1)
        best_i = np.take_along_axis(cat_i, order, axis=1)
    if ignore is not None:
        best_i = np.where(np.isneginf(best_s), -1, best_i)
    D = best_s if metric_type == METRIC_INNER_PRODUCT else -best_s
    return D, best_i


def check_topk(xb, xq, D, I, metric_type: int = METRIC_INNER_PRODUCT, rtol: float = 1e-5,
               atol: float = 0.0, ignore=None, D_ref=None, I_ref=None, what: str = ""):
    """Rigorous "exact top-k up to ties within tolerance" check against float64 arithmetic on
    the same inputs (SURVEY §7 'tie semantics'). Verifies, per query:
      1. every returned id is valid, unique and not the ignored id; padding only if k > rows;
      2. the returned score equals the float64 score of the RETURNED id within tol;
      3. returned scores are ordered (desc for IP, asc for L2) up to tol;
      4. no returned id is worse than the true k-th best by more than tol — i.e. the id set is
         a valid top-k; ids may differ from the float64 ranking only inside a tolerance tie.
    tol = rtol * max(|score|, |q||x|-scale of the row) + atol. Returns the number of positions
    whose id differs from the float64 reference (all of them certified ties)."""
    xb = np.asarray(xb)
    xq = np.asarray(xq)
    D = np.asarray(D, dtype=np.float64)
    I = np.asarray(I, dtype=np.int64)
    nq, k = I.shape
    if D_ref is None or I_ref is None:
        D_ref, I_ref = exact_topk_f64(xb, xq, k, metric_type, ignore=ignore)
    n_valid = xb.shape[0] - (0 if ignore is None else 1)
    n_diff = 0
    xb64 = xb.astype(np.float64)
    for q in range(nq):
        ids = I[q]
        kk = min(k, max(n_valid, 0)) if ignore is not None else min(k, xb.shape[0])
        good = ids[:kk]
        assert (good >= 0).all() and (good < xb.shape[0]).all(), f"{what}: q{q} invalid ids {ids}"
        assert len(set(good.tolist())) == kk, f"{what}: q{q} duplicate ids {ids}"
        assert (ids[kk:] == -1).all(), f"{what}: q{q} padding ids wrong {ids}"
        if ignore is not None:
            assert ignore[q] not in good.tolist(), f"{what}: q{q} returned the ignored id"
        if kk == 0:
            continue
        qv = xq[q].astype(np.float64)
        rows = xb64[good]
        ip = rows @ qv
        if metric_type == METRIC_INNER_PRODUCT:
            true = ip
            sign = 1.0
        else:
            true = (qv**2).sum() + (rows**2).sum(1) - 2 * ip
            sign = -1.0
        scale = np.sqrt((qv**2).sum()) * np.sqrt((rows**2).sum(1))
        tol = rtol * np.maximum(np.abs(true), scale) + atol
        got = D[q, :kk]
        assert (np.abs(got - true) <= tol).all(), (
            f"{what}: q{q} score mismatch got {got} true {true} tol {tol}")
        assert (sign * np.diff(got) <= 2 * tol[1:]).all(), f"{what}: q{q} not ordered: {got}"
        kth = D_ref[q, kk - 1]
        assert (sign * (true - kth) >= -2 * tol).all(), (
            f"{what}: q{q} id set is not a top-{kk}: true {true} kth {kth} ids {good} ref {I_ref[q]}")
        n_diff += int((good != I_ref[q, :kk]).sum())
    return n_diff


def recall_at_k(I_ref: np.ndarray, I: np.ndarray) -> float:
    hits = 0
    for a, b in zip(I_ref, I):
        hits += len(set(a[a >= 0].tolist()) & set(b[b >= 0].tolist()))
    denom = int((I_ref >= 0).sum())
    return hits / max(denom, 1)
