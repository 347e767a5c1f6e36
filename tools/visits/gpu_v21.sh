#!/bin/bash
# Round-2 visit 21: which of the pair kernel's three hand-offs (if any) should be a commit signal?  Kernel alone + in-step.
set -u
O=gpurun_out; mkdir -p $O
for h in arrive commit_p commit_ds commit_r1 commit; do
  echo "### DTA_PAIR_HANDOFF=$h"; DTA_PAIR_HANDOFF=$h timeout 120 python tools/attn_bench.py --case tau2 --iters 10
done > $O/pair_handoff_each.log 2>&1
grep -E "###|bwd_ms" $O/pair_handoff_each.log | cut -c1-330
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate"
for h in arrive commit_r1 arrive commit_r1; do
  DTA_PAIR_HANDOFF=$h python bench.py $F > $O/bench_ab5_${h}_$RANDOM.json 2> /dev/null; echo "$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab5_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), round(r['ms_per_launch'],2), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
