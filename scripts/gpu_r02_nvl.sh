#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/nvlink_bytes.py > $OUT/r02_nvlink_tp2.txt 2> $OUT/r02_nvlink_tp2.err
echo rc=$?; grep -v "^\[" $OUT/r02_nvlink_tp2.txt | tail -5; tail -2 $OUT/r02_nvlink_tp2.err
