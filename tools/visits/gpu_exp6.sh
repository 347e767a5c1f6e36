#!/bin/bash
set -u
O=gpurun_out/exp6.log; mkdir -p gpurun_out; : > $O
export DTA_K_ORDER_BUCKET=64
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 6 2>&1 | tail -1 >> $O; }
run DTA_BWD_L2PF=0
run DTA_BWD_L2PF=2
run DTA_BWD_L2PF=3
run DTA_BWD_L2PF=4
run DTA_BWD_L2PF=0
run DTA_BWD_L2PF=0 DTA_DEBUG_FLAGS=1
echo "### chain 0 / 3" >> $O
DTA_BWD_L2PF=0 timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 >> $O
DTA_BWD_L2PF=3 timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 >> $O
echo "### trace L2PF=3" >> $O
DTA_BWD_L2PF=3 timeout 200 python tools/bwd_trace.py 2>&1 | tail -4 >> $O
echo "### trace no reds" >> $O
DTA_DEBUG_FLAGS=1 timeout 200 python tools/bwd_trace.py 2>&1 | tail -4 >> $O
cat $O
