#!/bin/bash
set -u
O=gpurun_out/s4g.log; mkdir -p gpurun_out; : > $O
DTA_FWD_WARPS=16 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4 >> $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 8 2>&1 | tail -1 | cut -c60-260 >> $O; }
for rep in 1 2; do run DTA_FWD_WARPS=8; run DTA_FWD_WARPS=16 DTA_FWD_POLY=2; run DTA_FWD_WARPS=16 DTA_FWD_POLY=0; run DTA_FWD_WARPS=16 DTA_FWD_POLY=4; done
echo "### chain 8 / 16" >> $O
DTA_FWD_WARPS=8 timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 | cut -c60-260 >> $O
DTA_FWD_WARPS=16 timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 | cut -c60-260 >> $O
cat $O
