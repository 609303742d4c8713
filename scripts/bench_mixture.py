"""Generation / copy mixture (retriever_generator.py:391-404, forward) at BART sizes: the fused one-pass
kernel against the reference's own sequence of torch ops on the same GPU."""
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
print(json.dumps({"config": f"copy mixture B={B} T={T} V={V} S={S}", "ours_ms": ms_ours, "torch_ops_ms": ms_ref,
                  "speedup": ms_ref / ms_ours, "hbm_gbs_algorithmic": bytes_alg / ms_ours / 1e6,
                  "hbm_frac": bytes_alg / ms_ours / 1e6 / peaks["hbm_gbs"]}))
