#!/bin/bash
set -u
O=gpurun_out/exp8.log; mkdir -p gpurun_out; : > $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3 >> $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 6 2>&1 | tail -1 >> $O; }
run DTA_BWD_POLY=0
run DTA_BWD_POLY=2
run DTA_BWD_POLY=4
run DTA_BWD_POLY=0 DTA_BWD_L2PF=2
run DTA_BWD_POLY=0
echo "### chain" >> $O
timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 >> $O
echo "### trace" >> $O
timeout 200 python tools/bwd_trace.py 2>&1 | tail -4 >> $O
cat $O
