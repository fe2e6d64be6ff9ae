#!/bin/bash
# tests + bench A/B (PDL on/off) + ncu capture of the 7B GEMV kernels
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 400 $OUT/bench_$TAG.err
L2B_NO_PDL=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_nopdl.json 2> $OUT/bench_${TAG}_nopdl.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|attention_kernel' \
    -s 2270 -c 12 -f -o $OUT/prof_7b_$TAG python bench.py --workload llama2-7B --positions 4 --steps 1 --warmup 3 \
    --also none --no-cpu-baseline > $OUT/ncu_7b_$TAG.log 2>&1
tail -3 $OUT/ncu_7b_$TAG.log
ls -la $OUT | tail -8
