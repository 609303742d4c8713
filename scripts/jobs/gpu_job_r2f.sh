#!/bin/bash
O=gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/tools/debug_dist.py > $O/r2f_debug.log 2>&1
echo "debug rc=$?" >> $O/r2f_debug.log
tail -60 $O/r2f_debug.log
