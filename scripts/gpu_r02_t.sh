#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for MC in 256 128 64; do
  L2B_ATTN_MIN_CHUNK=$MC timeout 300 python bench.py --steps 3 --warmup 3 --also none --no-cpu-baseline > $OUT/r02t_bench_7b_mc$MC.json 2> $OUT/r02t_bench_7b_mc$MC.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02t_bench_7b_mc$MC.json")); print("MIN_CHUNK=$MC 7B", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), round(d["kernels"]["attention"]["ms"]*1e3,2), d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e)
PY
done
