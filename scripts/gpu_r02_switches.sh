#!/bin/bash
# every documented run-time switch still gives parity (a subset of the suite under each)
OUT=gpurun_out; mkdir -p $OUT
K="stories15m_logits_and_tokens or synthetic_models or argmax_and_generate or prefill_equals"
for ENVV in "L2B_NO_GRAPH=1" "L2B_NO_PDL=1" "L2B_ATTN=3pass" "L2B_GEMV_BIG=ldg" "L2B_FUSE=0" "L2B_TMA_STAGES=3" "L2B_TIME_STEPS=1" "L2B_PREFILL_BATCH=0"; do
  env $ENVV timeout 600 python -m pytest tests/test_gpu_transformer.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "$K" > $OUT/r02s_$ENVV.log 2>&1
  echo "$ENVV rc=$? $(tail -1 $OUT/r02s_$ENVV.log)"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  L2B_TP=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
     scripts/tp_check.py --shape 512,1376,3,8,8,-1024,96 --steps 12 > $OUT/r02s_tp_nccl.log 2>&1
  echo "L2B_TP=nccl rc=$? $(grep TP_CHECK $OUT/r02s_tp_nccl.log)"
fi
