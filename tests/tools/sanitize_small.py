"""Small end-to-end pass over every kernel for `compute-sanitizer --tool memcheck` (one tool per
gpurun call, B200_PROFILING.md): K0 (vector + scalar), K1 pair / 1-CTA / SIMT, K2, K3 (certified and
fallback), K4, K5, ragged shapes and both metrics."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m
from oracle import mips_oracle as o

rng = np.random.default_rng(0)
for d, n, nq, k in ((768, 1500, 260, 8), (100, 777, 37, 5), (1024, 900, 130, 16), (64, 300, 5, 33)):
    xb = rng.standard_normal((n, d), dtype=np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    xb[n // 2: n // 2 + 20] = xb[3]           # duplicates: certificate failures on the fp32 bank
    xq[:3] = xb[3] * 2
    for metric in (0, 1):
        for dtype, algos in (("bf16", ("tc2", "tc", "simt")), ("fp32", ("tcx", "simt"))):
            idx = m.B200FlatIndex(d, metric, dtype=dtype)
            idx.add(xb[: n // 2])
            idx.add(torch.from_numpy(xb[n // 2:]).cuda())
            ign = torch.arange(nq, dtype=torch.int64).cuda() % n
            for algo in algos:
                if algo == "tc" and d > 768:
                    continue
                r = idx.search_ex(torch.from_numpy(xq), k, ignore_ids=ign, algo=algo,
                                  want=("scores", "ids", "cosine", "doc_prob", "memory_bias"), L=7)
                torch.cuda.synchronize()
                stored = xb if dtype == "fp32" else o.bf16_round(xb)
                qq = xq if dtype == "fp32" else o.bf16_round(xq)
                o.check_topk(stored, qq, r["scores"].cpu().numpy(), r["ids"].cpu().numpy(), metric, rtol=1e-4,
                             ignore=ign.cpu().numpy(), what=f"{dtype} {algo} d={d} m={metric}")
            rows = idx.gather_rows(r["ids"])
            mt = m.retriever_metrics(r["ids"], torch.arange(n).cuda() % 7, torch.zeros(nq, dtype=torch.int64).cuda(),
                                     torch.ones(nq).cuda())
            idx.close()
store = m.MemoryTokenStore(rng.integers(0, 100, (50, 16)).astype(np.int32), lengths=np.full(50, 9), pad_id=1)
store.gather(torch.tensor([[0, 49, -1]], dtype=torch.int64).cuda())
torch.cuda.synchronize()
print("SANITIZE_SCRIPT_OK")
