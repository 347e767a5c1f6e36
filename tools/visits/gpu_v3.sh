#!/bin/bash
# Round-2 visit 3: do the hand-off waits cost their wake-up latency?  Polling waits (fwd flag 8, DTA_BWD_SPIN=1) vs parked.
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/fwd_trace.py > $O/fwd_trace_spin.log 2>&1; grep -E "###|period|step 2[01]" $O/fwd_trace_spin.log | cut -c1-420
for cfg in "1 0" "9 0" "1 1" "9 1"; do
  set -- $cfg
  for c in tau2 chain; do
    echo "### DTA_FWD_FLAGS=$1 DTA_BWD_SPIN=$2 $c"; DTA_FWD_FLAGS=$1 DTA_BWD_SPIN=$2 timeout 120 python tools/attn_bench.py --case $c --iters 10
  done
done > $O/spin_ab.log 2>&1
grep -E "###|fwd_ms" $O/spin_ab.log | cut -c1-330
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "logprob or vocab" > $O/pytest_logprob_v3.log 2>&1; echo "pytest logprob exit $?"; tail -3 $O/pytest_logprob_v3.log
