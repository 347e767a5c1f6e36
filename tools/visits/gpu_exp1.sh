#!/bin/bash
# experiments: UMMA operand-path rates, backward timeline, exp2 FMA-offload share, L2 prefetch, k-tile order
set -u
mkdir -p gpurun_out
O=gpurun_out/exp1.log
: > $O
tools/bin/umma_bench >> $O 2>&1
python tools/bwd_trace.py >> $O 2>&1
run() { echo "### $*" >> $O; env "$@" python tools/attn_bench.py --case tau2 --iters 4 2>&1 | tail -1 >> $O; }
run DTA_FWD_POLY=0 DTA_BWD_POLY=0
run DTA_FWD_POLY=1 DTA_BWD_POLY=1
run DTA_FWD_POLY=2 DTA_BWD_POLY=2
run DTA_FWD_POLY=3 DTA_BWD_POLY=3
run DTA_FWD_POLY=4 DTA_BWD_POLY=4
run DTA_FWD_POLY=0 DTA_BWD_POLY=0 DTA_DEBUG_FLAGS=16
run DTA_FWD_POLY=0 DTA_BWD_POLY=0 DTA_K_ORDER_BUCKET=64
run DTA_FWD_POLY=0 DTA_BWD_POLY=0 DTA_K_ORDER_BUCKET=100000
run DTA_FWD_POLY=0 DTA_BWD_POLY=0 DTA_K_ORDER_BUCKET=64 DTA_DEBUG_FLAGS=16
run DTA_FWD_POLY=0 DTA_BWD_POLY=0 DTA_DEBUG_FLAGS=1
echo "### chain" >> $O
DTA_FWD_POLY=0 DTA_BWD_POLY=0 python tools/attn_bench.py --case chain --iters 4 2>&1 | tail -1 >> $O
DTA_FWD_POLY=2 DTA_BWD_POLY=2 python tools/attn_bench.py --case chain --iters 4 2>&1 | tail -1 >> $O
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_exp1.log 2>&1; echo "pytest exit $?" >> $O
tail -3 gpurun_out/pytest_gpu_exp1.log >> $O
cat $O
