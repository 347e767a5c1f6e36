#!/bin/bash
# session-4 visit C: packed-fp32x2 forward softmax: parity + FMA-offload share sweep; UMMA interference microbenchmark
set -u
O=gpurun_out/s4c.log; mkdir -p gpurun_out; : > $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3 >> $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 6 2>&1 | tail -1 >> $O; }
run DTA_FWD_POLY=0
run DTA_FWD_POLY=2
run DTA_FWD_POLY=4
run DTA_FWD_POLY=6
run DTA_FWD_POLY=8
echo "### chain poly 2 / 4" >> $O
DTA_FWD_POLY=2 timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 >> $O
DTA_FWD_POLY=4 timeout 200 python tools/attn_bench.py --case chain --iters 6 2>&1 | tail -1 >> $O
echo "### umma_interf" >> $O
timeout 120 tools/bin/umma_interf >> $O 2>&1
cat $O
