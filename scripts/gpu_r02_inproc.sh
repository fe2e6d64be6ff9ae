#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; N=${1:-2}
timeout 400 python bench.py --in-process $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r02n_bench_inproc$N.json 2> $OUT/r02n_bench_inproc$N.err
echo "rc=$?"; tail -3 $OUT/r02n_bench_inproc$N.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/r02n_bench_inproc$N.json")); print("in-process N=$N 7B", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e: print("failed", e)
PY
