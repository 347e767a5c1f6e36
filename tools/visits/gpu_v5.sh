#!/bin/bash
# Round-2 visit 5: GPU suite (no -x), then the in-step A/B of the attention kernels: round-1 build vs current, same box.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_gpu.log; grep -E "^parity|full-model" $O/pytest_gpu.log | cut -c1-600
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong"
for i in 1 2; do
DTA_LIB_PATH=$PWD/tools/bin/libdta_r1attn.so python bench.py $F > $O/bench_ab_r1attn_$i.json 2> $O/bench_ab_r1attn_$i.err; echo "bench r1attn exit $?"
python bench.py $F > $O/bench_ab_cur_$i.json 2> $O/bench_ab_cur_$i.err; echo "bench cur exit $?"
done
DTA_FWD_FLAGS=0 python bench.py $F > $O/bench_ab_cur_fwd0.json 2> $O/bench_ab_cur_fwd0.err
DTA_BWD_PAIR=auto python bench.py $F > $O/bench_ab_cur_pair.json 2> $O/bench_ab_cur_pair.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), round(r['ms_per_launch'],2), 'fwd', round(r['attn_fwd_tflops']), round(r['attn_fwd_ms_per_launch'],2), d['clocks']['sm_mhz'], 'mem', round(d['peak_mem_gb'],1))
    except Exception as e: print(f, 'ERR', e)
PY
