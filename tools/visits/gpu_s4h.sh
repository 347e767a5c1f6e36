#!/bin/bash
set -u
O=gpurun_out/s4h.log; mkdir -p gpurun_out; : > $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case chain --iters 8 2>&1 | tail -1 | cut -c60-200 >> $O; }
run DTA_FWD_WARPS=8
run DTA_FWD_WARPS=16
run DTA_FWD_WARPS=16 DTA_FWD_DEBUG=1
run DTA_FWD_WARPS=16 DTA_FWD_DEBUG=2
run DTA_FWD_WARPS=16 DTA_FWD_DEBUG=4
run DTA_FWD_WARPS=16 DTA_FWD_DEBUG=7
cat $O
