#!/bin/bash
# round 2, call A: the refactored 1-GPU path — parity tests, then a short bench (1 GPU)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/r02a_gpu.txt 2>&1
free -g >> $OUT/r02a_gpu.txt; nproc >> $OUT/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    --deselect tests/test_gpu_tp.py > $OUT/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02a_pytest.log
tail -25 $OUT/r02a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02a_bench.json 2> $OUT/r02a_bench.err
echo "bench rc=$?"; tail -3 $OUT/r02a_bench.err; head -c 1500 $OUT/r02a_bench.json
