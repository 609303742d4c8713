#!/bin/bash
# round-2 job d: pooled admission threshold — correctness, A/B against MIPS_TC2_POOL=0, ncu of the C5 shard shape
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_robust.py -m gpu -q -x 2>&1 | tail -8 > $O/r2d_tests.log
rm -f $O/r2d_shapes.jsonl
for pool in 1 0; do
for cfg in "250000 1024 32 bf16 c5shard" "2000000 1024 32 bf16 c5full" "10000000 256 8 bf16 nq256" "1250000 1024 8 bf16 c3shard8" "10000000 1024 8 bf16 c3" "10000000 1024 64 bf16 c3k64" "250000 1024 8 bf16 small_k8"; do
  set -- $cfg
  MIPS_TC2_POOL=$pool python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --tag $5_pool$pool >> $O/r2d_shapes.jsonl 2>> $O/r2d_shapes.err
done
done
python scripts/prof_shape.py --rows 1000000 --nq 256 --k 8 --dtype fp32 --tag c2 >> $O/r2d_shapes.jsonl 2>> $O/r2d_shapes.err
ncu --set full --clock-control none --import-source on -k regex:search_tc2 --launch-skip 2 --launch-count 1 -o $O/r2d_c5shard -f python scripts/prof_shape.py --rows 250000 --nq 1024 --k 32 --ncu > $O/r2d_ncu_c5.log 2>&1
tail -3 $O/r2d_tests.log; cat $O/r2d_shapes.jsonl | cut -c1-250
