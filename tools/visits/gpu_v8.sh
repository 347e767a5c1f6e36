#!/bin/bash
# Round-2 visit 8: activation recomputation (tests, 1.7B test) and the --model 1.7b bench line.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_engine.py -m gpu -q -s > $O/pytest_engine_v8.log 2>&1; echo "pytest engine exit $?"; grep -E "1.7B|passed|failed|Error|assert" $O/pytest_engine_v8.log | cut -c1-700 | tail -12
python bench.py --model 1.7b --no-cpu-baseline --no-gpu-reference --no-strong > $O/bench_n1_1p7b.json 2> $O/bench_n1_1p7b.err; echo "bench 1.7b exit $?"; tail -3 $O/bench_n1_1p7b.err | cut -c1-400
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_1p7b.json')); c=d['config']; r=d['roofline']
print(d['value'], d['ms_per_step'], 'ranges', c['ranges_per_step'], 'recompute', c['activation_recompute'], 'mem', d['peak_mem_gb'], 'bwd', r['achieved'], 'fwd', r['attn_fwd_tflops'], 'share', r['attn_share_of_step'], 'dense', d['dense_baseline'])
PY
