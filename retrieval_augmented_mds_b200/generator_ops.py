"""Training-path operators at the END of the retrieval marginalisation, with gradients:

    copy_attention   sotasum/decoder_own.py:102-134,158-160 — the copy decoder's ONE-head cross attention over the
                     k*L memory tokens with the per-document logit `beta * mips_scores[b, doc] + beta_bias`
                     (attention_bias = memory_bias, retriever_generator.py:188-192), one softmax over all tokens
    copy_mixture     sotasum/retriever_generator.py:391-404 — log(gen_gate * softmax(logits)
                     + scatter_add(copy_probs) + 1e-7)

Both are `torch.autograd.Function`s over libmips_b200 kernels (k6_mixture.cuh, k7_attention.cuh). The two GEMMs of
the attention (q.k^T and p.v, and their four gradient GEMMs) are plain library GEMMs (`torch.bmm`); everything
elementwise / row-wise between them — bias, mask, softmax, its backward and the reduction that carries the
gradient to the retriever's document scores — is one fused kernel per direction. CUDA only: no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------ softmax
class _BiasedSoftmax(torch.autograd.Function):
    """P = softmax(scores + beta * doc_scores[b, s // mem_len] + beta_bias + mask[b, s]) over s."""

    @staticmethod
    def forward(ctx, scores, doc_scores, beta, beta_bias, mask, mem_len):
        B, T, S = scores.shape
        dev = scores.device
        scores = _f32c(scores)
        ds = None if doc_scores is None else _f32c(doc_scores)
        mk = None if mask is None else _f32c(mask)
        n_docs = 0 if ds is None else ds.shape[1]
        probs = torch.empty_like(scores)
        # trainable beta / beta_bias live on the device: the kernel reads them there (no .item() sync per step)
        on_dev = isinstance(beta, torch.Tensor) or isinstance(beta_bias, torch.Tensor)
        beta_t = beta.detach().reshape(-1)[:1].to(dev, torch.float32) if isinstance(beta, torch.Tensor) \
            else torch.full((1,), float(beta), device=dev)
        bias_t = beta_bias.detach().reshape(-1)[:1].to(dev, torch.float32) if isinstance(beta_bias, torch.Tensor) \
            else torch.full((1,), float(beta_bias), device=dev)
        beta_dev = torch.cat([beta_t, bias_t]) if on_dev else None
        with torch.cuda.device(dev):
            check(_lib.lib().mips_copy_attention_softmax_fwd(
                _ptr(scores), _ptr(ds), n_docs, int(mem_len), 0.0 if on_dev else float(beta),
                0.0 if on_dev else float(beta_bias), _ptr(beta_dev), _ptr(mk), B, T, S, _ptr(probs), _stream(dev)))
        ctx.save_for_backward(probs, ds if ds is not None else torch.empty(0, device=dev), beta_t)
        ctx.mem_len, ctx.has_docs = int(mem_len), ds is not None
        ctx.beta_is_tensor = isinstance(beta, torch.Tensor)
        ctx.bias_is_tensor = isinstance(beta_bias, torch.Tensor)
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        probs, ds, beta = ctx.saved_tensors
        B, T, S = probs.shape
        dev = probs.device
        dprobs = _f32c(dprobs)
        dscores = torch.empty_like(probs)
        G = torch.zeros_like(ds) if ctx.has_docs else None
        n_docs = ds.shape[1] if ctx.has_docs else 0
        with torch.cuda.device(dev):
            check(_lib.lib().mips_copy_attention_softmax_bwd(
                _ptr(probs), _ptr(dprobs), n_docs, ctx.mem_len, B, T, S, _ptr(dscores), _ptr(G), _stream(dev)))
        d_docs = d_beta = d_bias = None
        if ctx.has_docs:
            d_docs = beta * G                                  # -> mips_scores -> the query encoder (retriever)
            if ctx.beta_is_tensor:
                d_beta = (G * ds).sum().reshape(1)
            if ctx.bias_is_tensor:
                d_bias = G.sum().reshape(1)                    # ~0: a constant logit does not move a softmax
        return dscores, d_docs, d_beta, d_bias, None, None


def biased_softmax(scores: torch.Tensor, doc_scores: Optional[torch.Tensor], beta=1.0, beta_bias=0.0,
                   mask: Optional[torch.Tensor] = None, mem_len: int = 1) -> torch.Tensor:
    """scores [B, T, S]; doc_scores [B, n_docs] with token s of batch b belonging to document s // mem_len
    (mem_len = 1 and n_docs = S gives a general per-token bias = the reference's `attention_bias`); `mask`
    additive [B, S] (0 / finfo.min, what `_expand_mask` yields for one target position)."""
    if not scores.is_cuda:
        raise ValueError("biased_softmax runs on the GPU: pass CUDA tensors (no CPU compute path)")
    if scores.dim() != 3:
        raise ValueError("scores must be [B, T, S]")
    B, T, S = scores.shape
    if doc_scores is not None and (doc_scores.dim() != 2 or doc_scores.shape[0] != B or
                                   doc_scores.shape[1] * int(mem_len) < S):
        raise ValueError("doc_scores must be [B, n_docs] with n_docs * mem_len >= S")
    if mask is not None and tuple(mask.shape) != (B, S):
        raise ValueError("mask must be additive [B, S]")
    return _BiasedSoftmax.apply(scores, doc_scores, beta, beta_bias, mask, mem_len)


def copy_attention(query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor,
                   doc_scores: Optional[torch.Tensor], mem_len: int, beta=1.0, beta_bias=0.0,
                   attention_mask: Optional[torch.Tensor] = None):
    """The copy decoder's cross attention (decoder_own.py:102-134,158-160; one head): query_states [B, T, D] already
    projected and scaled (:72), key_states / value_states [B, S, D] with S = k * mem_len, doc_scores = `mips_scores`
    [B, k]; attention_mask additive [B, S] or the reference's [B, 1, T, S] (one slice is used: `_expand_mask`
    repeats the same row for every target position). Returns (attn_output [B, T, D], attn_weights [B, T, S]) —
    the second is the alignment that `copy_probs = copy_gate * alignment_weight` (:538) consumes. Differentiable in
    every tensor argument, including doc_scores (the retriever's learning signal), beta and beta_bias."""
    if attention_mask is not None and attention_mask.dim() == 4:
        attention_mask = attention_mask[:, 0, 0, :]
    scores = torch.bmm(query_states, key_states.transpose(1, 2))            # library GEMM (decoder_own.py:108)
    probs = biased_softmax(scores, doc_scores, beta, beta_bias, attention_mask, mem_len)
    return torch.bmm(probs, value_states), probs                            # library GEMM (decoder_own.py:160)


# ------------------------------------------------------------------------------------------------ mixture
class _CopyMixture(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, gen_gate, copy_probs, copy_seq, eps):
        B, T, V = logits.shape
        S = copy_probs.shape[2]
        dev = logits.device
        logits, gate, cp = _f32c(logits), _f32c(gen_gate).view(-1), _f32c(copy_probs)
        seq = copy_seq.to(torch.int64).contiguous()
        out = torch.empty((B, T, V), dtype=torch.float32, device=dev)
        stats = torch.empty((B * T, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().mips_copy_mixture_fwd(_ptr(logits), _ptr(gate), _ptr(cp), _ptr(seq), B * T, T, V, S,
                                                  float(eps), _ptr(out), _ptr(stats), _stream(dev)))
        ctx.save_for_backward(logits, gate, seq, out, stats)
        ctx.S, ctx.gate_shape = S, gen_gate.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        logits, gate, seq, out, stats = ctx.saved_tensors
        B, T, V = logits.shape
        dev = logits.device
        dout = _f32c(dout)
        dlogits = torch.empty_like(logits)
        dgate = torch.empty((B * T,), dtype=torch.float32, device=dev)
        dcopy = torch.empty((B, T, ctx.S), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().mips_copy_mixture_bwd(_ptr(logits), _ptr(out), _ptr(dout), _ptr(gate), _ptr(stats), _ptr(seq),
                                                  B * T, T, V, ctx.S, _ptr(dlogits), _ptr(dgate), _ptr(dcopy),
                                                  _stream(dev)))
        return dlogits, dgate.view(ctx.gate_shape), dcopy, None, None


def copy_mixture(logits: torch.Tensor, gen_gate: torch.Tensor, copy_probs: torch.Tensor, copy_seq: torch.Tensor,
                 eps: float = 1e-7) -> torch.Tensor:
    """log(gen_gate * softmax(logits) + scatter_add(copy_probs at copy_seq) + eps) in one pass over the logits
    (retriever_generator.py:391-404): logits [B, T, V], gen_gate [B, T, 1], copy_probs [B, T, S], copy_seq int64
    [B, S] -> [B, T, V] fp32, all CUDA. Differentiable in logits, gen_gate and copy_probs (one fused backward pass)."""
    if not (logits.is_cuda and gen_gate.is_cuda and copy_probs.is_cuda and copy_seq.is_cuda):
        raise ValueError("copy_mixture runs on the GPU: pass CUDA tensors (no CPU compute path)")
    if logits.dim() != 3 or copy_probs.dim() != 3 or copy_seq.dim() != 2:
        raise ValueError("expected logits [B, T, V], copy_probs [B, T, S], copy_seq [B, S]")
    B, T, V = logits.shape
    S = copy_probs.shape[2]
    if copy_probs.shape[:2] != (B, T) or copy_seq.shape != (B, S) or gen_gate.numel() != B * T:
        raise ValueError("inconsistent shapes")
    return _CopyMixture.apply(logits, gen_gate, copy_probs, copy_seq, eps)
