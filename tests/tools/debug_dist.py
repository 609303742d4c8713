"""2-rank walk through the sharded paths with progress markers (run under torchrun and `timeout`)."""
import faulthandler
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import retrieval_augmented_mds_b200 as m
from oracle import mips_oracle as o

faulthandler.dump_traceback_later(70, exit=True)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)


def mark(s):
    print(f"[r{rank} {time.time() % 1000:.2f}] {s}", flush=True)


dist.init_process_group("nccl", device_id=dev)
mark("pg up")
rng = np.random.default_rng(5)
n, d, nq, k = 40001, 256, 200, 8
xb = o.bf16_round(rng.standard_normal((n, d), dtype=np.float32))
xq = o.bf16_round(rng.standard_normal((nq, d), dtype=np.float32))
rows = m.shard_range(n, rank, world)
idx = m.B200FlatIndex(d, 0, dtype="bf16", device=dev)
sh = m.ShardedFlatIndex(idx, dist.group.WORLD)
sh.add_local(xb[rows.start:rows.stop])
mark("bank built")
D_ref, I_ref = o.exact_topk_f64(xb, xq, k)
c = sh.comm()
mark("native comm up")
xq_t = torch.from_numpy(xq).to(dev)
r = sh.search(xq_t, k)
torch.cuda.synchronize()
mark(f"native search ok={np.array_equal(r['ids'].cpu().numpy(), I_ref)}")
B = nq // world
mine = slice(rank * B, (rank + 1) * B)
r = sh.search_dp(xq_t[mine], k)
torch.cuda.synchronize()
mark(f"native dp search ok={np.array_equal(r['ids'].cpu().numpy(), I_ref[mine])}")
sh_t = m.ShardedFlatIndex(idx, dist.group.WORLD, exchange="torch")
sh_t.counts = sh.counts
r = sh_t.search(xq_t, k)
torch.cuda.synchronize()
mark(f"torch search ok={np.array_equal(r['ids'].cpu().numpy(), I_ref)}")
r = sh_t.search_dp(xq_t[mine], k)
torch.cuda.synchronize()
mark(f"torch dp search ok={np.array_equal(r['ids'].cpu().numpy(), I_ref[mine])}")
g = sh.capture(nq, k)
mark("captured")
out = g.replay(xq_t)
torch.cuda.synchronize()
mark(f"replay ok={np.array_equal(out['ids'].cpu().numpy(), I_ref)}")
gd = sh.capture(B, k, dp=True)
out = gd.replay(xq_t[mine])
torch.cuda.synchronize()
mark(f"dp replay ok={np.array_equal(out['ids'].cpu().numpy(), I_ref[mine])}")
shp = m.ShardedFlatIndex(idx, dist.group.WORLD, exchange="p2p")
shp.counts = sh.counts
for it in range(3):
    a = shp.search(xq_t, k)
    torch.cuda.synchronize()
    mark(f"p2p search {it} ok={np.array_equal(a['ids'].cpu().numpy(), I_ref)}")
shp.check_exchange()
shp.close()
sh.close()
mark("closed")
dist.destroy_process_group()
mark("done")
