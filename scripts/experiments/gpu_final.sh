#!/bin/bash
# Round-end style validation on ONE GPU: smoke, gpu tests, default bench, ncu launch list + full captures.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/gpu_$TAG.txt
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -1 $OUT/smoke_$TAG.log
timeout 400 python -m pytest tests -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; tail -2 $OUT/pytest_gpu_$TAG.log
timeout 400 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 300 $OUT/bench_$TAG.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 330 --csv \
    --log-file $OUT/launches_15m_$TAG.csv python bench.py --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/ncu_launch15_$TAG.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 330 --csv \
    --log-file $OUT/launches_7b_$TAG.csv python bench.py --workload llama2-7B --positions 8 --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/ncu_launch7b_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemv_tma_kernel|attention' \
    -s 2600 -c 12 -f -o $OUT/prof_7b_$TAG python bench.py --workload llama2-7B --positions 4 --steps 1 --warmup 3 \
    --also none --no-cpu-baseline > $OUT/ncu_7b_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|attention' \
    -s 330 -c 33 -f -o $OUT/prof_15m_$TAG python bench.py --positions 32 --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/ncu_15m_$TAG.log 2>&1
ls $OUT | grep $TAG | tr '\n' ' '
