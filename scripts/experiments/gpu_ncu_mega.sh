#!/bin/bash
TAG=${1:-mega}
OUT=gpurun_out
L2B_MEGA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mega_step_kernel' \
    -s 30 -c 2 -f -o $OUT/prof_mega15_$TAG python bench.py --workload stories15M --positions 32 --steps 1 --warmup 3 \
    --also none --no-cpu-baseline > $OUT/ncu_mega15_$TAG.log 2>&1
tail -3 $OUT/ncu_mega15_$TAG.log
