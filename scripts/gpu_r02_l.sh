#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
L2B_FUSE=0 timeout 120 python scripts/trace_step.py stories15M > $OUT/r02l_trace_15m_chain.txt 2> $OUT/r02l_trace_15m.err
sed -n 7,12p $OUT/r02l_trace_15m_chain.txt
timeout 120 python scripts/trace_step.py llama2-7B > $OUT/r02l_trace_7b.txt 2> $OUT/r02l_trace_7b.err
sed -n 7,12p $OUT/r02l_trace_7b.txt; tail -8 $OUT/r02l_trace_7b.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
