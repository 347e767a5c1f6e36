#!/bin/bash
set -u
O=gpurun_out/s4l.log; mkdir -p gpurun_out; : > $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 8 2>&1 | tail -1 | cut -c60-260 >> $O; }
for rep in 1 2 3; do run DTA_LIB_PATH=$PWD/tools/bin/libdta_presplit.so; run X=1; done
cat $O
