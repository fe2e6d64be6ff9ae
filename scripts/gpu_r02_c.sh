#!/bin/bash
# round 2, call C (1 GPU): fused small-model kernels + producer warp freed from the prologue barriers
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    --deselect tests/test_gpu_tp.py -x > $OUT/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02c_pytest.log; tail -12 $OUT/r02c_pytest.log
for F in 3 1 0; do
  L2B_FUSE=$F timeout 300 python bench.py --workload stories15M --also stories110M --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02c_bench_small_fuse$F.json 2> $OUT/r02c_bench_small_fuse$F.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02c_bench_small_fuse$F.json"))
    print("FUSE=$F 15M", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
    r=d["also"]["stories110M"]; print("FUSE=$F 110M", round(r["value"]), "e2e", round(r["e2e"]["value"]), {k: round(v["ms"]*1e3,2) for k,v in r["kernels"].items()})
except Exception as e: print("bench small failed", e); print(open("$OUT/r02c_bench_small_fuse$F.err").read()[-1500:])
PY
done
for R in 2 8; do
  L2B_ATTN_R=$R timeout 300 python bench.py --workload stories15M --also stories110M --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02c_bench_small_R$R.json 2> $OUT/r02c_bench_small_R$R.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02c_bench_small_R$R.json"))
    print("R=$R 15M", round(d["value"]), "110M", round(d["also"]["stories110M"]["value"]))
except Exception as e: print("bench small R failed", e)
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --also none > $OUT/r02c_bench_7b.json 2> $OUT/r02c_bench_7b.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/r02c_bench_7b.json")); print("7B", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), d["roofline"]["frac"], d["whole_step"]["frac_of_peak"], {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
except Exception as e: print("bench 7b failed", e); print(open("$OUT/r02c_bench_7b.err").read()[-1500:])
PY
timeout 200 python scripts/trace_step.py llama2-7B > $OUT/r02c_trace_7b_1gpu.txt 2> $OUT/r02c_trace_1gpu.err
tail -9 $OUT/r02c_trace_7b_1gpu.txt
