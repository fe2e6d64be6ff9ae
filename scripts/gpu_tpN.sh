#!/bin/bash
# usage: gpu_tpN.sh TAG N [small shape for tp_check]
TAG=$1; N=$2
OUT=gpurun_out
mkdir -p $OUT
SHAPE=${3:-1024,2752,2,16,8,-2048,48}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   scripts/tp_check.py --shape $SHAPE --steps 12 > $OUT/tpcheck_$TAG.log 2>&1
grep -E "TP_CHECK" $OUT/tpcheck_$TAG.log || tail -5 $OUT/tpcheck_$TAG.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/bench_$TAG.json") if l.startswith("{")][-1]
    print("N=$N", round(d["value"],1), "tok/s e2e", round(d["e2e"]["value"],1), {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("bench FAILED", e); print(open("$OUT/bench_$TAG.err").read()[-800:])
PY
