#!/bin/bash
# Round-2 visit 22: dS^T hand-off as a commit signal, in the step (same box, alternating).
set -u
O=gpurun_out; mkdir -p $O
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate"
i=0
for h in arrive commit_ds commit_p_ds arrive commit_ds commit_p_ds; do
  i=$((i+1)); DTA_PAIR_HANDOFF=$h python bench.py $F > $O/bench_ab6_${i}_${h}.json 2> /dev/null; echo "$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab6_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), round(r['ms_per_launch'],2), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
DTA_PAIR_HANDOFF=commit_ds python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "pair or benchmarked" 2>&1 | tail -3
