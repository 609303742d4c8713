#!/bin/bash
O=gpurun_out
export BENCH_WATCHDOG_S=100
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -5 > $O/r2n_tests.log
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2n_bench_n2.json 2> $O/r2n_bench_n2.err
echo "rc=$?"; tail -2 $O/r2n_tests.log; python - <<'PY'
import json
d = [json.loads(l) for l in open('gpurun_out/r2n_bench_n2.json') if l.startswith('{')][-1]
print(round(d['value']), d['ms_per_step'], d['roofline']['k1_ms_avg'], d['roofline']['step_minus_k1_ms'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['parity']['ok'])
PY
grep -v "NCCL INFO" $O/r2n_bench_n2.err | grep -i "error\|Timeout\|File" | head
