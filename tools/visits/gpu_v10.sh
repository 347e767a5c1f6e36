#!/bin/bash
# Round-2 visit 10: in-step A/B (same box): head-major launch order of the CTA-pair backward; forward exp2 split.
set -u
O=gpurun_out; mkdir -p $O
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong"
run() { name=$1; shift; env "$@" python bench.py $F > $O/bench_ab3_$name.json 2> $O/bench_ab3_$name.err; echo "$name exit $?"; }
run base1 X=1
run bwdhm1 DTA_BWD_HEAD_MAJOR=1
run base2 X=1
run bwdhm2 DTA_BWD_HEAD_MAJOR=1
run fwdpoly0 DTA_FWD_POLY=0
run fwdpoly4 DTA_FWD_POLY=4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab3_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), round(r['ms_per_launch'],2), 'fwd', round(r['attn_fwd_tflops']), round(r['attn_fwd_ms_per_launch'],2), d['clocks']['sm_mhz'], r['kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
