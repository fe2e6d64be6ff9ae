#!/bin/bash
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --workload llama2-7B --also none --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_tma.json 2> $OUT/bench_${TAG}_tma.err
tail -c 300 $OUT/bench_${TAG}_tma.err
L2B_NO_PDL=1 timeout 900 python bench.py --workload llama2-7B --also none --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_tma_nopdl.json 2> $OUT/bench_${TAG}_tma_nopdl.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemv_tma_kernel|attention' \
    -s 2500 -c 10 -f -o $OUT/prof_7b_$TAG python bench.py --workload llama2-7B --positions 4 --steps 1 --warmup 3 \
    --also none --no-cpu-baseline > $OUT/ncu_7b_$TAG.log 2>&1
tail -2 $OUT/ncu_7b_$TAG.log
