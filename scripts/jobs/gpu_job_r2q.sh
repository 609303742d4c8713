#!/bin/bash
# round-2 final verification (1 GPU): full GPU tests, smoke, bench both arms, launch list + one full ncu capture of K1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2q_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2q_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2q_bench_n1.json 2> $O/r2q_bench_n1.err
BENCH_REF_BUDGET_S=25 timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2q_bench_ref.json 2> $O/r2q_bench_ref.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2q_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > $O/r2q_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_tc2 --launch-skip 3 --launch-count 1 -o $O/r2q_k1 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > $O/r2q_ncu_full.log 2>&1
tail -3 $O/r2q_tests.log; tail -2 $O/r2q_smoke.log; cut -c1-700 $O/r2q_bench_n1.json; cut -c1-900 $O/r2q_bench_ref.json; tail -2 $O/r2q_bench_ref.err; ls -la $O/r2q_k1.ncu-rep $O/r2q_launches.csv
