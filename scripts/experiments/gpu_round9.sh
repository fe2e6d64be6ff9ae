#!/bin/bash
TAG=${1:-r01k}
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
run flash7b llama2-7B L2B_X=1
run pass3_7b llama2-7B L2B_ATTN=3pass
run flash15 stories15M L2B_X=1
run pass3_15 stories15M L2B_ATTN=3pass
run flash110 stories110M L2B_X=1
L2B_TRACE=1 timeout 200 python scripts/trace_step.py llama2-7B > $OUT/trace_7b_$TAG.txt 2>&1; tail -8 $OUT/trace_7b_$TAG.txt
