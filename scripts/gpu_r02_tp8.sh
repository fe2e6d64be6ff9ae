#!/bin/bash
# round 2, 8-GPU call: tensor-parallel parity at 4 and 8 ranks, llama2-7B bench at N=8 and N=4, trace at N=8
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02f}
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider \
    -k "4-512 or 8-1024 or 8-shape3" > $OUT/${TAG}_pytest_tp.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_tp.log; tail -6 $OUT/${TAG}_pytest_tp.log
for W in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29513 \
     bench.py --gpus $W --steps 5 --warmup 3 > $OUT/${TAG}_bench_tp$W.json 2> $OUT/${TAG}_bench_tp$W.err
  echo "bench tp$W rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/${TAG}_bench_tp$W.json") if l.startswith("{")][-1]
    print("N=$W 7B", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), d.get("parity"), {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()}, d["clocks"])
except Exception as e:
    print("bench FAILED", e); print(open("$OUT/${TAG}_bench_tp$W.err").read()[-1500:])
PY
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 \
   scripts/trace_step.py llama2-7B > $OUT/${TAG}_trace_7b_tp8.txt 2> $OUT/${TAG}_trace_tp8.err
echo "trace tp8 rc=$?"; tail -9 $OUT/${TAG}_trace_7b_tp8.txt
