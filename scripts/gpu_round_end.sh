#!/bin/bash
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/gpu_$TAG.txt
timeout 400 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 300 $OUT/bench_$TAG.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 330 --csv \
    --log-file $OUT/launches_15m_$TAG.csv python bench.py --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/ncu_launch15_$TAG.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 330 --csv \
    --log-file $OUT/launches_7b_$TAG.csv python bench.py --workload llama2-7B --positions 8 --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/ncu_launch7b_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'gemv_tma_kernel|attention' \
    -s 2600 -c 6 -f -o $OUT/prof_7b_$TAG python bench.py --workload llama2-7B --positions 4 --steps 1 --warmup 3 \
    --also none --no-cpu-baseline > $OUT/ncu_7b_$TAG.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|attention' \
    -s 330 -c 7 -f -o $OUT/prof_15m_$TAG python bench.py --positions 32 --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/ncu_15m_$TAG.log 2>&1
du -sh $OUT; ls -la $OUT | grep $TAG | awk '{print $5, $9}' | tr '\n' ' '
