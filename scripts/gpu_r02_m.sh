#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_transformer.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    -k "not 7b_full" > $OUT/r02m_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02m_pytest.log; tail -5 $OUT/r02m_pytest.log
L2B_FUSE=0 timeout 120 python scripts/trace_step.py stories15M > $OUT/r02m_trace_15m_chain.txt 2> $OUT/r02m_trace_15m.err
sed -n 7,12p $OUT/r02m_trace_15m_chain.txt
timeout 120 python scripts/trace_step.py llama2-7B > $OUT/r02m_trace_7b.txt 2> $OUT/r02m_trace_7b.err
sed -n 7,9p $OUT/r02m_trace_7b.txt; tail -8 $OUT/r02m_trace_7b.txt | head -7
for F in 1 0; do
L2B_FUSE=$F timeout 300 python bench.py --workload stories15M --also stories110M --no-cpu-baseline --steps 5 --warmup 3 > $OUT/r02m_bench_small_f$F.json 2> $OUT/r02m_bench_small_f$F.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/r02m_bench_small_f$F.json")); print("FUSE=$F 15M", round(d["value"]), "e2e", round(d["e2e"]["value"]), "110M", round(d["also"]["stories110M"]["value"]), d["clocks"])
except Exception as e: print("failed", e)
PY
done
