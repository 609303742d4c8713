#!/bin/bash
# round-2 job b: full GPU test suite, shape timings, ncu captures of the small-shard / exact / ridge shapes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2b_tests.log
for cfg in "250000 1024 32 bf16 c5shard" "2000000 1024 32 bf16 c5full" "1000000 256 8 fp32 c2" "10000000 256 8 bf16 nq256" "1250000 1024 8 bf16 c3shard8" "10000000 16 5 bf16 c4"; do
  set -- $cfg
  python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --tag $5 >> $O/r2b_shapes.jsonl 2>> $O/r2b_shapes.err
done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:search_tc2 --launch-skip 2 --launch-count 1 -o $O/r2b_c5shard -f python scripts/prof_shape.py --rows 250000 --nq 1024 --k 32 --ncu > $O/r2b_ncu_c5.log 2>&1
$NCU -k regex:search_tc2 --launch-skip 2 --launch-count 1 -o $O/r2b_nq256 -f python scripts/prof_shape.py --rows 10000000 --nq 256 --k 8 --ncu > $O/r2b_ncu_nq256.log 2>&1
$NCU -k regex:"search_tc2|merge_topk|rerank" --launch-skip 6 --launch-count 3 -o $O/r2b_c2 -f python scripts/prof_shape.py --rows 1000000 --nq 256 --k 8 --dtype fp32 --ncu > $O/r2b_ncu_c2.log 2>&1
tail -5 $O/r2b_tests.log; cat $O/r2b_shapes.jsonl; ls -la $O/*.ncu-rep
