#!/bin/bash
set -u
O=gpurun_out/s4d.log; mkdir -p gpurun_out; : > $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 8 2>&1 | tail -1 | cut -c60-260 >> $O; }
for rep in 1 2; do for v in 0 1 2 3 5 7; do run DTA_FWD_VARIANT=$v DTA_FWD_POLY=2; done; done
run DTA_FWD_VARIANT=3 DTA_FWD_POLY=0
run DTA_FWD_VARIANT=3 DTA_FWD_POLY=4
DTA_FWD_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2 >> $O
cat $O
