#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r02r_bench.json 2> $OUT/r02r_bench.err
echo "bench rc=$? lines=$(wc -l < $OUT/r02r_bench.json)"; tail -2 $OUT/r02r_bench.err
python - <<PY
import json
d=json.load(open("$OUT/r02r_bench.json"))
print("7B", round(d["value"],2), "e2e", d["e2e"], "traffic", d["roofline"]["traffic"], d["roofline"]["frac"], d["clocks"])
for w,r in d["also"].items(): print(w, round(r["value"]), r["e2e"])
print({k: d[k] for k in ("metric","unit","n_gpus","steps","warmup","higher_is_better","scaling","dtype","gpu_launches")})
PY
