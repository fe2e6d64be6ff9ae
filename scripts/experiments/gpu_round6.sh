#!/bin/bash
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload llama2-7B --also none --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$name.json"))
    print("$name", round(d["value"],1), "tok/s", round(d["whole_step"]["achieved_gbs_per_gpu"],1), "GB/s", {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("$name FAILED", e)
PY
}
run half_auto L2B_X=1
run full5_1cta L2B_TMA_STAGES=5
run st3_2cta L2B_TMA_CTAS=2
run st2_2cta L2B_TMA_STAGES=2 L2B_TMA_CTAS=2
