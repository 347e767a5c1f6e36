#!/bin/bash
# Round-2 visit 30: fp32 dQ hand-over (attention backward -> q/k-norm backward): parity, then in-step A/B.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -k "qknorm or engine or fused_trunk or recompute or 28_layer or memory_bounded" > $O/pytest_v30.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_v30.log | cut -c1-300
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate"
i=0
for h in 0 1 0 1 0 1; do
  i=$((i+1)); DTA_FUSE_QK_ATTN=$h python bench.py $F > $O/bench_ab8_${i}_fuse$h.json 2> /dev/null; echo "fuse=$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab8_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'], 'mem', round(d['peak_mem_gb'],1), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
