#!/bin/bash
# Round-2 visit 43: TMA-store dK / dV epilogue of the single-CTA backward -- kernel tests and the bushy end of the sweep.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $O/pytest_kernels_v43.log 2>&1
echo "pytest exit $?"; tail -4 $O/pytest_kernels_v43.log
timeout 600 python tools/attn_bench.py --case bushy --iters 7 > $O/attn_bushy_v43.jsonl 2> $O/attn_bushy_v43.err
echo "sweep exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/attn_bushy_v43.jsonl'):
    d = json.loads(l)
    print(d['case'], 'fwd', d['fwd_tflops'], 'bwd', d['bwd_tflops'], 'bwd_ms', d['bwd_ms'])
PY
