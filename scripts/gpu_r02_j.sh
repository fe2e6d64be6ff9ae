#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python scripts/trace_cluster.py stories15M > $OUT/r02j_trace_cluster.txt 2> $OUT/r02j_trace_cluster.err
cat $OUT/r02j_trace_cluster.txt; tail -3 $OUT/r02j_trace_cluster.err
