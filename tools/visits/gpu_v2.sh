#!/bin/bash
# Round-2 visit 2: forward-kernel pipeline variants (DTA_FWD_FLAGS: 1 head-major order, 2 split P, 4 split S) and the
# head-major order of the backward (DTA_BWD_HEAD_MAJOR), A/B on one box; parity suite with the new defaults.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tree_attention or chain_equals or constant_values or leaf_order or schedules_agree" > $O/pytest_attn_v2.log 2>&1; echo "pytest attn exit $?"; tail -2 $O/pytest_attn_v2.log
timeout 300 python tools/fwd_trace.py > $O/fwd_trace.log 2>&1; grep -E "###|period" $O/fwd_trace.log | cut -c1-300
for f in 0 1 2 4 6 7; do
  for c in tau2 chain; do
    echo "### DTA_FWD_FLAGS=$f DTA_BWD_HEAD_MAJOR=0 $c"; DTA_FWD_FLAGS=$f DTA_BWD_HEAD_MAJOR=0 timeout 120 python tools/attn_bench.py --case $c --iters 10
  done
done > $O/fwd_flags_ab.log 2>&1
echo "### DTA_FWD_FLAGS=7 DTA_BWD_HEAD_MAJOR=1 tau2" >> $O/fwd_flags_ab.log; DTA_BWD_HEAD_MAJOR=1 timeout 120 python tools/attn_bench.py --case tau2 --iters 10 >> $O/fwd_flags_ab.log 2>&1
echo "### DTA_FWD_FLAGS=7 DTA_BWD_HEAD_MAJOR=1 DTA_BWD_PAIR=1 tau2" >> $O/fwd_flags_ab.log; DTA_BWD_PAIR=1 DTA_BWD_HEAD_MAJOR=1 timeout 120 python tools/attn_bench.py --case tau2 --iters 10 >> $O/fwd_flags_ab.log 2>&1
grep -E "###|fwd_ms" $O/fwd_flags_ab.log | cut -c1-330
