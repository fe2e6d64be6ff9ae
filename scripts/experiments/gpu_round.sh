#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list + one full capture of the top kernels.
# Usage: bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/gpu_$TAG.txt
nproc >> $OUT/gpu_$TAG.txt; free -g | head -2 >> $OUT/gpu_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 600 $OUT/bench_$TAG.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 330 --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --also none --no-cpu-baseline \
      > $OUT/ncu_launch_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|attention_kernel' \
      -s 66 -c 33 -f -o $OUT/prof_15m_$TAG python bench.py --steps 1 --warmup 3 --also none --no-cpu-baseline \
      > $OUT/ncu_full_$TAG.log 2>&1
fi
ls -la $OUT
