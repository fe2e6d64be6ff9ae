#!/bin/bash
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu_$TAG.log 2>&1
tail -4 $OUT/pytest_gpu_$TAG.log
run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 240 python bench.py --workload $wl --also none --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$name.json"))
    print("$name", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), "whole", round(d["whole_step"]["achieved_gbs_per_gpu"],1), "GB/s")
except Exception as e:
    print("$name FAILED", e); print(open("$OUT/bench_${TAG}_$name.err").read()[-600:])
PY
}
run mega15 stories15M L2B_MEGA=1
run graph15 stories15M L2B_MEGA=0


run mega7b llama2-7B L2B_MEGA=1
run graph7b llama2-7B L2B_MEGA=0
