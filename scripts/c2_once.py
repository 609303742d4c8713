"""C2 (1M x 768 fp32, 256 queries, k=8) a few times: the command behind the ncu launch list of the
exact tensor-core search (filter + re-rank + certificate)."""
import sys
import torch
sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m

dev = torch.device("cuda:0")
n, d, nq, k = 1_000_000, 768, 256, 8
idx = m.B200FlatIndex(d, 0, dtype="fp32", capacity=n)
gen = torch.Generator(device=dev).manual_seed(1)
for s in range(0, n, 250_000):
    idx.add(torch.randn((250_000, d), generator=gen, device=dev))
xq = torch.randn((nq, d), generator=gen, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    r = idx.search_ex(xq, k)
torch.cuda.synchronize()
print(idx.last_algo, idx.fallback_queries(), r["ids"][0].tolist())
