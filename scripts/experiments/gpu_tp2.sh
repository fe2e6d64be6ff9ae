#!/bin/bash
TAG=${1:-tp2}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$TAG.txt
nvidia-smi topo -m >> $OUT/gpus_$TAG.txt 2>&1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   scripts/tp_check.py --steps 16 > $OUT/tpcheck_$TAG.log 2>&1
grep -E "TP_CHECK|Error|error" $OUT/tpcheck_$TAG.log | head -5
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   scripts/tp_check.py --shape 4096,11008,2,32,32,-32000,64 --steps 6 > $OUT/tpcheck7b_$TAG.log 2>&1
grep -E "TP_CHECK|Error|error" $OUT/tpcheck7b_$TAG.log | head -5
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 600 $OUT/bench_$TAG.err
L2B_TP=nccl timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_${TAG}_nccl.json 2> $OUT/bench_${TAG}_nccl.err
python - <<PY
import json
for name in ("$OUT/bench_$TAG.json", "$OUT/bench_${TAG}_nccl.json"):
    try:
        d=[json.loads(l) for l in open(name) if l.startswith("{")][-1]
        print(name, round(d["value"],1), "tok/s", {k: round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(name, "FAILED", e)
PY
