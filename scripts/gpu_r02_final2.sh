#!/bin/bash
# round 2, last 1-GPU pass on the final tree: full parity suite, default bench, ncu launch lists + captures
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider \
    --deselect tests/test_gpu_tp.py > $OUT/${TAG}_pytest_final.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_final.log; tail -4 $OUT/${TAG}_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_r02_final.sh $TAG
