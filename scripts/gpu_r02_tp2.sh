#!/bin/bash
# round 2, call B (2 GPUs): tensor-parallel parity (multi-process + in-process + dead peer), TP bench + trace,
# 1-GPU trace of the refactored prologue, classifier TPR experiment.  Every stage has its own timeout.
OUT=gpurun_out; mkdir -p $OUT; N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > $OUT/r02b_pytest_tp.log 2>&1
echo "pytest rc=$?" >> $OUT/r02b_pytest_tp.log; tail -15 $OUT/r02b_pytest_tp.log
if grep -q "failed" $OUT/r02b_pytest_tp.log; then echo "TP TESTS FAILED - still running the bench for timing info"; fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus $N --steps 3 --warmup 3 > $OUT/r02b_bench_tp$N.json 2> $OUT/r02b_bench_tp$N.err
echo "bench tp$N rc=$?"; tail -3 $OUT/r02b_bench_tp$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
   scripts/trace_step.py llama2-7B > $OUT/r02b_trace_7b_tp$N.txt 2> $OUT/r02b_trace_tp$N.err
echo "trace tp$N rc=$?"; tail -12 $OUT/r02b_trace_7b_tp$N.txt
timeout 200 python scripts/trace_step.py llama2-7B > $OUT/r02b_trace_7b_1gpu.txt 2> $OUT/r02b_trace_1gpu.err
echo "trace 1gpu rc=$?"; tail -10 $OUT/r02b_trace_7b_1gpu.txt
for T in 0 2; do
  L2B_TPR_TILES=$T timeout 200 python bench.py --workload stories15M --also none --no-cpu-baseline --steps 5 --warmup 3 \
     > $OUT/r02b_bench15_tpr$T.json 2> $OUT/r02b_bench15_tpr$T.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02b_bench15_tpr$T.json")); print("TPR_TILES=$T 15M", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms"]*1e3,2) for k,v in d["kernels"].items()})
except Exception as e: print("bench15 failed", e)
PY
done
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/r02b_bench_tp$N.json") if l.startswith("{")][-1]
    print("N=$N 7B", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), d.get("parity"), {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("bench FAILED", e); print(open("$OUT/r02b_bench_tp$N.err").read()[-1500:])
PY
