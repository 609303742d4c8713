#!/bin/bash
# round-2 final regression (1 GPU): full GPU test suite + smoke on the final code state
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2u_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2u_smoke.log 2>&1
tail -3 $O/r2u_tests.log; tail -2 $O/r2u_smoke.log
