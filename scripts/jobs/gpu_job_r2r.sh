#!/bin/bash
O=gpurun_out
export BENCH_WATCHDOG_S=120
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > $O/r2r_bench_n1.json 2> $O/r2r_bench_n1.err
echo "rc=$?"; python - <<'PY'
import json
d = [json.loads(l) for l in open('gpurun_out/r2r_bench_n1.json') if l.startswith('{')][-1]
r = d['roofline']
print(round(d['value']), d['ms_per_step'], 'k1', r['k1_ms_avg'], r['k1_launches_timed'], r['frac'], r['frac_of_burst_peak'], r['step_minus_k1_ms'], r['k1_timed_in'][:60], 'e2e', round(d['e2e']['value']), d['parity']['ok'], d['gpu_launches'])
PY
tail -3 $O/r2r_bench_n1.err
