#!/bin/bash
# round 2, call H (1 GPU): the all-layers cluster kernel for small models: parity suite, then A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_ops.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    -x -k "not 7b and not 110m" > $OUT/r02h_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02h_pytest.log; tail -15 $OUT/r02h_pytest.log
for CL in 16 8 0; do
  L2B_CLUSTER=$CL timeout 300 python bench.py --workload stories15M --also none --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02h_bench15_cl$CL.json 2> $OUT/r02h_bench15_cl$CL.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02h_bench15_cl$CL.json"))
    print("CLUSTER=$CL 15M", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
except Exception as e: print("bench failed", e); print(open("$OUT/r02h_bench15_cl$CL.err").read()[-1200:])
PY
done
L2B_CLUSTER=0 L2B_FUSE=0 timeout 120 python scripts/trace_step.py stories15M > $OUT/r02h_trace_15m_chain.txt 2> $OUT/r02h_trace_15m.err
head -14 $OUT/r02h_trace_15m_chain.txt
