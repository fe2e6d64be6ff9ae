#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for CFG in "1024 8388608" "512 67108864" "512 8388608"; do
  set -- $CFG
  L2B_BIG_MIN_N=$1 L2B_GEMV8_MIN_BYTES=$2 timeout 300 python bench.py --workload stories110M --also none --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02p_bench110_$1_$2.json 2> $OUT/r02p_bench110_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02p_bench110_$1_$2.json")); print("BIG_MIN_N=$1 MIN_BYTES=$2 110M", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e)
PY
done
