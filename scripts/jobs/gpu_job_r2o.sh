#!/bin/bash
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -70 > $O/r2o_tests.log; tail -3 $O/r2o_tests.log
grep -n "Error\|error\|assert\|line [0-9]* in" $O/r2o_tests.log | head -30
