#!/bin/bash
# the driver's own commands at N=2: both arms under torchrun with its step counts
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $OUT/r02u_ref_tp2.json 2> $OUT/r02u_ref_tp2.err
echo "reference arm rc=$? lines=$(grep -c '^{' $OUT/r02u_ref_tp2.json) secs=$(( $(date +%s) - T0 ))"; head -c 400 $OUT/r02u_ref_tp2.json; echo
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
   bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r02u_bench_tp2.json 2> $OUT/r02u_bench_tp2.err
echo "gpu arm rc=$? lines=$(grep -c '^{' $OUT/r02u_bench_tp2.json) secs=$(( $(date +%s) - T0 ))"
python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/r02u_bench_tp2.json") if l.startswith("{")][-1]
r=[json.loads(l) for l in open("$OUT/r02u_ref_tp2.json") if l.startswith("{")][-1]
print("N=2", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["parity"], d["steps"], d["warmup"], d["clocks"])
print("same config:", d["config"] == r["config"], "steps", r["steps"], r["warmup"], "ref value", r["value"])
PY
