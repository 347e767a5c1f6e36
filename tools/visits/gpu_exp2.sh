#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/exp2.log
: > $O
run() { echo "### $*" >> $O; env "$@" python tools/attn_bench.py --case tau2 --iters 6 2>&1 | tail -1 >> $O; }
export DTA_K_ORDER_BUCKET=64
for g in 0 2 4 6 8 12 16; do run DTA_BWD_REDGAP=$g; done
run DTA_BWD_REDGAP=4 DTA_BWD_POLY=2
run DTA_BWD_REDGAP=4 DTA_BWD_POLY=4
run DTA_BWD_REDGAP=8 DTA_BWD_POLY=2
run DTA_BWD_REDGAP=8 DTA_BWD_POLY=4
echo "### trace gap 0" >> $O; DTA_BWD_REDGAP=0 python tools/bwd_trace.py 2>&1 | tail -7 >> $O
echo "### trace gap 6" >> $O; DTA_BWD_REDGAP=6 python tools/bwd_trace.py 2>&1 | tail -7 >> $O
cat $O
