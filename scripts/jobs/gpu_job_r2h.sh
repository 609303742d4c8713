#!/bin/bash
# round-2 job h: new K2 (stage once, REDUX rounds) — full GPU tests, shapes, launch lists
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > $O/r2h_tests.log
rm -f $O/r2h_shapes.jsonl
for cfg in "250000 1024 32 bf16 c5shard" "1000000 256 8 fp32 c2" "1250000 1024 8 bf16 c3shard8" "10000000 16 5 bf16 c4" "10000000 128 64 bf16 nq128k64" "1000000 256 64 fp32 c2k64"; do
  set -- $cfg
  timeout 200 python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --tag $5 >> $O/r2h_shapes.jsonl 2>> $O/r2h_shapes.err
done
for cfg in "250000 1024 32 bf16 c5shard" "1250000 1024 8 bf16 c3shard8" "1000000 256 8 fp32 c2"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2h_launches_$5.csv python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --ncu > $O/r2h_ncu_$5.log 2>&1
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/r2h_launches_*.csv')):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ci = {n: i for i, n in enumerate(hdr)}
    print(f)
    for r in rows[-7:]:
        print('   ', r[ci['Kernel Name']][:60], r[ci['Metric Value']], r[ci['Metric Unit']])
PY
tail -4 $O/r2h_tests.log; cut -c1-260 $O/r2h_shapes.jsonl
