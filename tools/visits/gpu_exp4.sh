#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/exp4.log
: > $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3 >> $O
export DTA_K_ORDER_BUCKET=64
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 6 2>&1 | tail -1 >> $O; }
run DTA_BWD_POLY=0
run DTA_BWD_POLY=2
run DTA_BWD_POLY=3
echo "### chain" >> $O; timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 >> $O
echo "### trace" >> $O; timeout 200 python tools/bwd_trace.py 2>&1 | tail -7 >> $O
cat $O
