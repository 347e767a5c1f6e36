#!/bin/bash
# compute-sanitizer pass over small cases of every kernel family (the reference has no race /
# memory checking; this is ours).  memcheck: out-of-bounds and misaligned global / shared accesses; synccheck: invalid
# barrier usage.  racecheck is left out on purpose: it does not model mbarrier / async-proxy (TMA, tcgen05) ordering and
# reports the kernels' producer-consumer hand-offs as hazards.  NOT YET RUN (queued for round 2).  ~5 minutes.
set -u
O=gpurun_out; mkdir -p $O
SEL='fan_small or chain_129 or single_token or adversarial or logprob_golden or add_rmsnorm or swiglu or qknorm'
for tool in memcheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py -q -x -k "$SEL" > $O/sanitize_$tool.log 2>&1
  echo "compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitize_$tool.log | tail -3
done
