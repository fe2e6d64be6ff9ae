#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > $OUT/r02q_pytest_tp.log 2>&1
echo "pytest rc=$?" >> $OUT/r02q_pytest_tp.log; tail -4 $OUT/r02q_pytest_tp.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02q_smoke.log 2>&1; tail -2 $OUT/r02q_smoke.log
