#!/bin/bash
# round-2 job g: launch lists (per-kernel durations) of one search for the C5 shard, C3 shard-of-8 and C2 shapes
O=gpurun_out
for cfg in "250000 1024 32 bf16 c5shard" "1250000 1024 8 bf16 c3shard8" "1000000 256 8 fp32 c2" "10000000 1024 8 bf16 c3"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2g_launches_$5.csv python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --ncu > $O/r2g_ncu_$5.log 2>&1
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/r2g_launches_*.csv')):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ci = {n: i for i, n in enumerate(hdr)}
    body = rows[1:]
    print(f)
    for r in body[-14:]:
        print('   ', r[ci['Kernel Name']][:70], r[ci['Metric Value']], r[ci['Metric Unit']])
PY
