#!/bin/bash
# round-2 job k (8 GPUs): headline bench at N=8 (graph + native NCCL), the round-1 style step for comparison, C5 and C4-DP
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2k_bench_n8.json 2> $O/r2k_bench_n8.err
timeout 200 $TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-graph --exchange torch --no-parity > $O/r2k_bench_n8_torch.json 2> $O/r2k_bench_n8_torch.err
timeout 200 $TR --master-port 29523 scripts/bench_c5.py > $O/r2k_c5_n8.json 2> $O/r2k_c5_n8.err
timeout 200 $TR --master-port 29524 scripts/bench_c4_dp.py > $O/r2k_c4dp_n8.json 2> $O/r2k_c4dp_n8.err
timeout 200 $TR --master-port 29525 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2k_bench_n8_b.json 2> $O/r2k_bench_n8_b.err
for f in r2k_bench_n8 r2k_bench_n8_torch r2k_bench_n8_b; do python - $f <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open('gpurun_out/' + sys.argv[1] + '.json') if l.startswith('{')][-1]
    print(sys.argv[1], round(d['value']), d['ms_per_step'], d['roofline']['k1_ms_avg'], d['roofline']['step_minus_k1_ms'], round(d['e2e']['value']), d['config']['step'], d.get('parity'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
cat $O/r2k_c5_n8.json $O/r2k_c4dp_n8.json; grep -c "NCCL INFO" $O/r2k_bench_n8.err; grep "nranks" $O/r2k_bench_n8.err | head -3; tail -3 $O/r2k_c5_n8.err $O/r2k_c4dp_n8.err | cut -c1-300
