#!/bin/bash
# round 2, call K (1 GPU): full parity suite after the attention merge rewrite, then the full default bench
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    --deselect tests/test_gpu_tp.py > $OUT/r02k_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02k_pytest.log; tail -6 $OUT/r02k_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02k_bench.json 2> $OUT/r02k_bench.err
echo "bench rc=$?"; tail -2 $OUT/r02k_bench.err
python - <<PY
import json
d=json.load(open("$OUT/r02k_bench.json"))
print("7B", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "frac", round(d["roofline"]["frac"],4), round(d["whole_step"]["frac_of_peak"],4), d.get("prefill",{}).get("tokens_per_s"), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
for w,r in d["also"].items(): print(w, round(r["value"]), "e2e", round(r["e2e"]["value"]), r.get("prefill",{}).get("tokens_per_s"), {k: round(v["ms"]*1e3,2) for k,v in r["kernels"].items()})
print(d["cpu_baseline"]); print(d["clocks"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02k_bench_ref.json 2> $OUT/r02k_bench_ref.err
echo "ref rc=$?"; head -c 600 $OUT/r02k_bench_ref.json
