#!/bin/bash
# round 2 profile pass (1 GPU): ncu launch lists of the bench command + `--set full` captures of the hot kernels
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
python -c "import bench; print(bench.kernel_source_sha())" > $OUT/${TAG}_prof_sha.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/${TAG}_gpu.txt
# launch list, llama2-7B (the headline workload): 8 positions of one step, after the warm-up runs
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 340 --csv \
    --log-file $OUT/${TAG}_launches_7b.csv python bench.py --workload llama2-7B --positions 8 --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/${TAG}_ncu_launch7b.log 2>&1
echo "launch7b rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 330 --csv \
    --log-file $OUT/${TAG}_launches_15m.csv python bench.py --workload stories15M --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/${TAG}_ncu_launch15.log 2>&1
echo "launch15 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemv_tma_kernel|attention' \
    -s 3200 -c 6 -f -o $OUT/${TAG}_prof_7b python bench.py --workload llama2-7B --positions 4 --steps 1 --warmup 3 \
    --also none --no-cpu-baseline > $OUT/${TAG}_ncu_7b.log 2>&1
echo "full7b rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|attention|attn_wo|ffn' \
    -s 330 -c 7 -f -o $OUT/${TAG}_prof_15m python bench.py --workload stories15M --positions 32 --steps 1 --warmup 3 --also none --no-cpu-baseline \
    > $OUT/${TAG}_ncu_15m.log 2>&1
echo "full15 rc=$?"
du -sh $OUT; ls -la $OUT | grep ${TAG}_ | awk '{print $5, $9}' | tr '\n' ' '
