#!/bin/bash
O=gpurun_out
export BENCH_WATCHDOG_S=100
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2m_bench_n8.json 2> $O/r2m_bench_n8.err
echo "rc=$?"; cat $O/r2m_bench_n8.json | cut -c1-3000; grep -v "NCCL INFO" $O/r2m_bench_n8.err | grep -i "error\|Timeout\|File" | head -20
