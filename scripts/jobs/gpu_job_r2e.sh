#!/bin/bash
# round-2 job e (2 GPUs): the sharded paths — NCCL through the C ABI, data-parallel queries, graph capture, smoke, bench N=2
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -30 > $O/r2e_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2e_smoke.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2e_bench_n2.json 2> $O/r2e_bench_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph --exchange torch --no-parity > $O/r2e_bench_n2_torch.json 2> $O/r2e_bench_n2_torch.err
tail -5 $O/r2e_tests.log; tail -3 $O/r2e_smoke.log; cat $O/r2e_bench_n2.json | cut -c1-2600; grep -c "NCCL INFO" $O/r2e_bench_n2.err; grep -i "nranks\|error\|Traceback" $O/r2e_bench_n2.err | head -8; cat $O/r2e_bench_n2_torch.json | cut -c1-400
