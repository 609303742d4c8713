"""Generation / copy mixture (retriever_generator.py:391-404) and the score-biased copy attention
(decoder_own.py:102-134, 160-176) at the retriever-generator step's sizes (B=16, T=256, BART vocabulary, k*L = 2560
memory tokens, D=1024), forward and forward+backward: the fused kernels against the reference's own sequence of torch
ops (under autograd) on the same GPU."""
import json
import sys

import torch

sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m

dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
gen = torch.Generator(device=dev).manual_seed(0)
B, T, V, S = 16, 256, 50265, 2560
logits = torch.randn((B, T, V), generator=gen, device=dev)
gates = torch.softmax(torch.randn((B, T, 2), generator=gen, device=dev), -1)
gen_gate = gates[..., :1].contiguous()
copy_probs = gates[..., 1:] * torch.softmax(torch.randn((B, T, S), generator=gen, device=dev), -1)
copy_seq = torch.randint(0, V, (B, S), generator=gen, device=dev)
index = copy_seq.reshape(B, 1, -1).expand(-1, T, -1)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ref():
    probs = gen_gate * torch.softmax(logits, -1)
    probs.scatter_add_(-1, index, copy_probs)
    return torch.log(probs + 1e-7)


ms_ours = timed(lambda: m.copy_mixture(logits, gen_gate, copy_probs, copy_seq))
ms_ref = timed(ref)
bytes_alg = B * T * (8 * V + 12 * S)
print(json.dumps({"config": f"copy mixture forward B={B} T={T} V={V} S={S}", "ours_ms": ms_ours, "torch_ops_ms": ms_ref,
                  "speedup": ms_ref / ms_ours, "hbm_gbs_algorithmic": bytes_alg / ms_ours / 1e6,
                  "hbm_frac": bytes_alg / ms_ours / 1e6 / peaks["hbm_gbs"]}))

# forward + backward (training)
lg, gg, cp = logits.clone().requires_grad_(True), gen_gate.clone().requires_grad_(True), copy_probs.clone().requires_grad_(True)
w = torch.randn((B, T, V), generator=gen, device=dev)


def train_ours():
    m.copy_mixture(lg, gg, cp, copy_seq).backward(w)
    lg.grad = gg.grad = cp.grad = None


def train_ref():
    probs = gg * torch.softmax(lg, -1)
    probs = probs.scatter_add(-1, index, cp)
    torch.log(probs + 1e-7).backward(w)
    lg.grad = gg.grad = cp.grad = None


ms_ours, ms_ref = timed(train_ours, 10), timed(train_ref, 10)
bytes_alg = B * T * (8 * V + 12 * S) + B * T * (16 * V + 12 * S)
print(json.dumps({"config": f"copy mixture forward+backward B={B} T={T} V={V} S={S}", "ours_ms": ms_ours,
                  "torch_ops_ms": ms_ref, "speedup": ms_ref / ms_ours, "hbm_gbs_algorithmic": bytes_alg / ms_ours / 1e6,
                  "hbm_frac": bytes_alg / ms_ours / 1e6 / peaks["hbm_gbs"]}))

# score-biased copy attention: D = 1024, k = 5 documents of L = 512 tokens
D, K, L = 1024, 5, 512
q = (torch.randn((B, T, D), generator=gen, device=dev) * D ** -0.5).requires_grad_(True)
kk = torch.randn((B, S, D), generator=gen, device=dev).requires_grad_(True)
vv = torch.randn((B, S, D), generator=gen, device=dev).requires_grad_(True)
doc = torch.rand((B, K), generator=gen, device=dev).requires_grad_(True)
beta = torch.nn.Parameter(torch.ones(1, device=dev))
beta_bias = torch.nn.Parameter(torch.zeros(1, device=dev))
add_mask = torch.zeros((B, S), device=dev)
add_mask[:, -100:] = torch.finfo(torch.float32).min
wo, wp = torch.randn((B, T, D), generator=gen, device=dev), torch.randn((B, T, S), generator=gen, device=dev)
params = (q, kk, vv, doc, beta, beta_bias)


def attn_ours(train):
    out, p = m.copy_attention(q, kk, vv, doc, L, beta, beta_bias, add_mask)
    if train:
        ((out * wo).sum() + (p * wp).sum()).backward()
        for t in params:
            t.grad = None


def attn_ref(train):
    aw = torch.bmm(q, kk.transpose(1, 2))
    bias = doc.unsqueeze(-1).expand(-1, -1, L).reshape(B, -1)                       # retriever_generator.py:188-192
    aw = aw + (beta * bias.view(B, 1, -1) + beta_bias)                              # decoder_own.py:110-114
    aw = (aw.view(B, 1, T, S) + add_mask[:, None, None, :]).view(B, T, S)           # :123-132
    p = torch.softmax(aw, -1)
    out = torch.bmm(p, vv)
    if train:
        ((out * wo).sum() + (p * wp).sum()).backward()
        for t in params:
            t.grad = None


for train in (False, True):
    with torch.set_grad_enabled(train):
        a, b = timed(lambda: attn_ours(train), 10), timed(lambda: attn_ref(train), 10)
    print(json.dumps({"config": f"copy attention {'forward+backward' if train else 'forward'} B={B} T={T} S={S} D={D} "
                                f"(incl. the library GEMMs)", "ours_ms": a, "torch_ops_ms": b, "speedup": b / a}))
