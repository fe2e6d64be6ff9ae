#!/bin/bash
# round 2, call D1 (N GPUs): TP with the producer-free prologue: parity tests, bench, trace
OUT=gpurun_out; mkdir -p $OUT; N=${1:-2}; TAG=${2:-r02d}
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > $OUT/${TAG}_pytest_tp.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_tp.log; tail -6 $OUT/${TAG}_pytest_tp.log
for W in $N $3; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29513 \
     bench.py --gpus $W --steps 3 --warmup 3 > $OUT/${TAG}_bench_tp$W.json 2> $OUT/${TAG}_bench_tp$W.err
  echo "bench tp$W rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/${TAG}_bench_tp$W.json") if l.startswith("{")][-1]
    print("N=$W 7B", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), d.get("parity"), {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("bench FAILED", e); print(open("$OUT/${TAG}_bench_tp$W.err").read()[-1500:])
PY
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29515 \
     scripts/trace_step.py llama2-7B > $OUT/${TAG}_trace_7b_tp$W.txt 2> $OUT/${TAG}_trace_tp$W.err
  echo "trace tp$W rc=$?"; tail -9 $OUT/${TAG}_trace_7b_tp$W.txt
done
for PF in 1 2; do
  L2B_TMA_PREFILL=$PF timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_tp${N}_pf$PF.json 2> $OUT/${TAG}_bench_tp${N}_pf$PF.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/${TAG}_bench_tp${N}_pf$PF.json") if l.startswith("{")][-1]
    print("PREFILL=$PF N=$N 7B", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("bench FAILED", e)
PY
done
