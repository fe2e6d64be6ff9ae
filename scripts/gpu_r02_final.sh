#!/bin/bash
# round 2, final 1-GPU pass: default bench (both arms), then the ncu launch lists and --set full captures
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err
echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench_1gpu.err
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_1gpu.json"))
print("7B", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "frac", round(d["roofline"]["frac"],4), round(d["whole_step"]["frac_of_peak"],4), d.get("prefill",{}).get("tokens_per_s"), d["clocks"])
for w,r in d["also"].items(): print(w, round(r["value"]), "e2e", round(r["e2e"]["value"]))
PY
bash scripts/gpu_r02_profile.sh $TAG
