#!/bin/bash
# round 2, call D2 (1 GPU): full parity suite with the new entry points, ring pre-fill depth A/B on 7B,
# small-model fusion A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    --deselect tests/test_gpu_tp.py > $OUT/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02d_pytest.log; tail -8 $OUT/r02d_pytest.log
for PF in 0 2; do
  L2B_TMA_PREFILL=$PF timeout 300 python bench.py --steps 3 --warmup 3 --also none --no-cpu-baseline > $OUT/r02d_bench_7b_pf$PF.json 2> $OUT/r02d_bench_7b_pf$PF.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02d_bench_7b_pf$PF.json")); print("PREFILL=$PF 7B", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), round(d["whole_step"]["frac_of_peak"],4), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
except Exception as e: print("bench 7b failed", e); print(open("$OUT/r02d_bench_7b_pf$PF.err").read()[-800:])
PY
done
L2B_TMA_PREFILL=2 timeout 200 python scripts/trace_step.py llama2-7B > $OUT/r02d_trace_7b_pf2.txt 2> $OUT/r02d_trace_pf2.err
tail -9 $OUT/r02d_trace_7b_pf2.txt
for CFG in "0 4" "1 8" "3 8"; do
  set -- $CFG
  L2B_FUSE=$1 L2B_ATTN_R=$2 timeout 300 python bench.py --workload stories15M --also stories110M --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02d_bench_small_f$1_r$2.json 2> $OUT/r02d_bench_small_f$1_r$2.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02d_bench_small_f$1_r$2.json"))
    print("FUSE=$1 R=$2 15M", round(d["value"]), "e2e", round(d["e2e"]["value"]), "110M", round(d["also"]["stories110M"]["value"]), "e2e", round(d["also"]["stories110M"]["e2e"]["value"]))
except Exception as e: print("bench small failed", e)
PY
done
