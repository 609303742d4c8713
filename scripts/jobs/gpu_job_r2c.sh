#!/bin/bash
# round-2 job c: branch-free candidate extraction — correctness subset, shape timings, stage-size sweep at the HBM-bound ridge
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_robust.py -m gpu -q -x 2>&1 | tail -8 > $O/r2c_tests.log
rm -f $O/r2c_shapes.jsonl
for cfg in "250000 1024 32 bf16 c5shard" "2000000 1024 32 bf16 c5full" "1000000 256 8 fp32 c2" "10000000 256 8 bf16 nq256" "1250000 1024 8 bf16 c3shard8" "10000000 1024 8 bf16 c3" "10000000 128 8 bf16 nq128"; do
  set -- $cfg
  python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --tag $5 >> $O/r2c_shapes.jsonl 2>> $O/r2c_shapes.err
done
for sk in 2 4 6; do
  MIPS_TC2_SKCH=$sk python scripts/prof_shape.py --rows 10000000 --nq 256 --k 8 --tag nq256_skch$sk >> $O/r2c_shapes.jsonl 2>> $O/r2c_shapes.err
  MIPS_TC2_SKCH=$sk python scripts/prof_shape.py --rows 1000000 --nq 256 --k 8 --dtype fp32 --tag c2_skch$sk >> $O/r2c_shapes.jsonl 2>> $O/r2c_shapes.err
done
python scripts/prof_shape.py --rows 10000000 --nq 256 --k 8 --algo tc --tag nq256_tc >> $O/r2c_shapes.jsonl 2>> $O/r2c_shapes.err
tail -3 $O/r2c_tests.log; cat $O/r2c_shapes.jsonl | cut -c1-330
