#!/bin/bash
# First GPU visit of round 2: is the CTA-pair attention backward (DTA_BWD_PAIR=1, csrc/attn_bwd_pair.cu) faster on
# the bench shape, and which hand-off form should it use?  A/B on ONE box: kernel-level parity + tau2 timing, the
# hand-off latency bench, the variant sweep, the both-CTA timeline, the whole GPU suite with the pair kernel on, and
# the bench line with it off and with the intended production rule (auto).  Everything lands in gpurun_out/.  ~6 min.
set -u
O=gpurun_out; mkdir -p $O tools/bin
timeout 120 python tools/bwd_pair_check.py > $O/bwd_pair_check.log 2>&1; echo "pair check exit $?"; tail -2 $O/bwd_pair_check.log | cut -c1-600
# hand-off latency of the pair's barriers (arrive vs multicast commit, idle and under reduce traffic)
[ -x tools/bin/pair_handoff_bench ] || nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I dynamictreeattn_b200/csrc -I include \
    tools/pair_handoff_bench.cu dynamictreeattn_b200/csrc/dta_common.cu -lcuda -o tools/bin/pair_handoff_bench
for m in 0 1 2 3; do timeout 30 tools/bin/pair_handoff_bench $m; echo "handoff mode $m exit $?"; done > $O/pair_handoff.log 2>&1; cat $O/pair_handoff.log
# variant sweep on the bench shape; the commit-signal hand-off only if a non-leader commit proved legal above
COMMIT=0
if grep -q "handoff mode 3 exit 0" $O/pair_handoff.log && ! grep -q "mode 3.*TIMED OUT" $O/pair_handoff.log; then
  DTA_PAIR_HANDOFF=commit timeout 120 python tools/bwd_pair_check.py > $O/bwd_pair_check_commit.log 2>&1 && COMMIT=1
  echo "pair check (commit hand-off) ok=$COMMIT"; tail -1 $O/bwd_pair_check_commit.log | cut -c1-400
fi
PAIR_SWEEP_COMMIT=$COMMIT timeout 150 python tools/bwd_pair_check.py --sweep tau2 > $O/bwd_pair_sweep.log 2>&1; cut -c1-200 $O/bwd_pair_sweep.log
timeout 60 python tools/bwd_pair_trace.py > $O/bwd_pair_trace.log 2>&1; head -2 $O/bwd_pair_trace.log; grep period $O/bwd_pair_trace.log
DTA_BWD_PAIR=1 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_pair_on.log 2>&1; echo "pytest (pair on) exit $?"; tail -2 $O/pytest_gpu_pair_on.log
python bench.py --no-cpu-baseline > $O/bench_n1_pair_off.json 2> $O/bench_n1_pair_off.err; echo "bench (pair off) exit $?"; cut -c1-160 $O/bench_n1_pair_off.json
DTA_BWD_PAIR=auto python bench.py --no-cpu-baseline --no-dense-baseline > $O/bench_n1_pair_auto.json 2> $O/bench_n1_pair_auto.err; echo "bench (pair auto) exit $?"; cut -c1-160 $O/bench_n1_pair_auto.json
