#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/r2j_tests.log
timeout 200 python scripts/prof_shape.py --rows 10000000 --nq 1024 --k 100 --tag c3k100 > $O/r2j_shapes.jsonl 2> $O/r2j_shapes.err
tail -6 $O/r2j_tests.log; cut -c1-260 $O/r2j_shapes.jsonl; tail -3 $O/r2j_shapes.err
