#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_transformer.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "prefill" > $OUT/r02o_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02o_pytest.log; tail -5 $OUT/r02o_pytest.log
