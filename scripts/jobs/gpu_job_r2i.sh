#!/bin/bash
# round-2 job i: multi-pass k > 64, bounded kernels — full GPU tests + headline shape regression check
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/r2i_tests.log
rm -f $O/r2i_shapes.jsonl
for cfg in "10000000 1024 8 bf16 c3" "250000 1024 32 bf16 c5shard" "1000000 256 8 fp32 c2" "10000000 256 8 bf16 nq256" "10000000 1024 100 bf16 c3k100"; do
  set -- $cfg
  timeout 200 python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --tag $5 >> $O/r2i_shapes.jsonl 2>> $O/r2i_shapes.err
done
tail -6 $O/r2i_tests.log; cut -c1-260 $O/r2i_shapes.jsonl; tail -3 $O/r2i_shapes.err
