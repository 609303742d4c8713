#!/bin/bash
O=gpurun_out
export BENCH_WATCHDOG_S=120
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2s_bench_n2.json 2> $O/r2s_bench_n2.err
echo "rc=$?"; python - <<'PY'
import json
lines = [l for l in open('gpurun_out/r2s_bench_n2.json')]
print('stdout lines', len(lines))
d = [json.loads(l) for l in lines if l.startswith('{')][-1]
r = d['roofline']
print(round(d['value']), d['ms_per_step'], 'k1', r['k1_ms_avg'], r['k1_launches_timed'], r['frac_of_burst_peak'], r['step_minus_k1_ms'], 'e2e', round(d['e2e']['value']), d['parity']['ok'], d['gpu_launches'])
PY
grep -v "NCCL INFO" $O/r2s_bench_n2.err | grep -i "error\|Timeout\|Traceback" | head -5; grep -c "NCCL" $O/r2s_bench_n2.err
