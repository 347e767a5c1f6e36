#!/bin/bash
# Round-2 visit 40: persistent forward as the default -- full GPU suite, smoke, bench (N = 1), step profile, and the
# full config-5 sweep with both forward kernels side by side.
set -u
bash tools/gpu_final.sh
O=gpurun_out
timeout 900 python tools/attn_bench.py --case sweep --iters 5 > $O/attn_sweep_config5_persist.jsonl 2> $O/attn_sweep_config5_persist.err
echo "sweep exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/attn_sweep_config5_persist.jsonl'):
    d = json.loads(l)
    print(d['case'], 'default', d['fwd_tflops'], 'one-item', d['fwd_one_item_per_cta_tflops'], 'persistent', d['fwd_persistent_tflops'], 'bwd', d['bwd_tflops'])
PY
