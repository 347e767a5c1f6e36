#!/bin/bash
# Round-2 visit 6: in-step A/B of the forward launch order (template variant, hot loop unchanged) on one box, with the
# CTA-pair backward as the new default; then the full default bench line.
set -u
O=gpurun_out; mkdir -p $O
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong"
for i in 1 2; do
DTA_FWD_HEAD_MAJOR=0 python bench.py $F > $O/bench_ab2_hm0_$i.json 2> $O/bench_ab2_hm0_$i.err; echo "bench hm0 exit $?"
DTA_FWD_HEAD_MAJOR=1 python bench.py $F > $O/bench_ab2_hm1_$i.json 2> $O/bench_ab2_hm1_$i.err; echo "bench hm1 exit $?"
done
DTA_BWD_PAIR=0 python bench.py $F > $O/bench_ab2_pair0.json 2> $O/bench_ab2_pair0.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab2_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), round(r['ms_per_launch'],2), 'fwd', round(r['attn_fwd_tflops']), round(r['attn_fwd_ms_per_launch'],2), d['clocks']['sm_mhz'], r['kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench full exit $?"; cut -c1-300 $O/bench_n1.json
