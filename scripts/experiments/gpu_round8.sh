#!/bin/bash
TAG=${1:-r01j}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 240 python bench.py --workload $wl --also none --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$name.json"))
    print("$name", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), "whole", round(d["whole_step"]["achieved_gbs_per_gpu"],1), "GB/s", {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("$OUT/bench_${TAG}_$name.err").read()[-600:])
PY
}
run pf192 llama2-7B L2B_PF_KB=192
run pf0 llama2-7B L2B_PF_KB=0
run pf384 llama2-7B L2B_PF_KB=384
run pf96 llama2-7B L2B_PF_KB=96
