#!/bin/bash
O=gpurun_out
export BENCH_WATCHDOG_S=45
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-parity > $O/r2l_bench_n2.json 2> $O/r2l_bench_n2.err
echo "rc=$?"; cat $O/r2l_bench_n2.json | cut -c1-300; grep -v "NCCL INFO" $O/r2l_bench_n2.err | tail -60 | cut -c1-200
