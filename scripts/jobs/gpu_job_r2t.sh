#!/bin/bash
# round-2 job t: A/B of the K2 radix-cut threshold (512 vs 2048) on the shapes whose local merge stages 576 / 1184 candidates
O=gpurun_out
rm -f $O/r2t_shapes.jsonl
for cut in 512 2048; do
  for cfg in "250000 1024 32 bf16 c5shard" "1000000 256 8 fp32 c2"; do
    set -- $cfg
    MIPS_K2_CUT=$cut timeout 100 python scripts/prof_shape.py --rows $1 --nq $2 --k $3 --dtype $4 --tag $5_cut$cut >> $O/r2t_shapes.jsonl 2>> $O/r2t_shapes.err
  done
done
MIPS_K2_CUT=2048 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robust.py -m gpu -q -x 2>&1 | tail -3
cut -c1-200 $O/r2t_shapes.jsonl
