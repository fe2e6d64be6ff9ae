#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_transformer.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider \
    -x -k "not 7b and not 110m and not long_context and not batched" > $OUT/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02i_pytest.log; tail -4 $OUT/r02i_pytest.log
timeout 120 python scripts/trace_cluster.py stories15M > $OUT/r02i_trace_cluster.txt 2> $OUT/r02i_trace_cluster.err
cat $OUT/r02i_trace_cluster.txt; tail -3 $OUT/r02i_trace_cluster.err
for CL in 16 0; do
  L2B_CLUSTER=$CL timeout 300 python bench.py --workload stories15M --also none --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02i_bench15_cl$CL.json 2> $OUT/r02i_bench15_cl$CL.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02i_bench15_cl$CL.json"))
    print("CLUSTER=$CL 15M", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
except Exception as e: print("bench failed", e); print(open("$OUT/r02i_bench15_cl$CL.err").read()[-1200:])
PY
done
