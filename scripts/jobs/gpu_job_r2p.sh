#!/bin/bash
# round-2 job p: vectorised mixture kernels — correctness + timings
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_generator_ops.py tests/test_gpu_interchange.py -m gpu -q -x 2>&1 | tail -6 > $O/r2p_tests.log
timeout 200 python scripts/bench_mixture.py > $O/r2p_mixture.jsonl 2> $O/r2p_mixture.err
tail -3 $O/r2p_tests.log; cat $O/r2p_mixture.jsonl; tail -3 $O/r2p_mixture.err
