#!/bin/bash
# Round-2 visit 12: fused lm_head (N1) parity, engine integration, micro-bench, in-step A/B.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -k "fused_lm_head or fused_head or launches_counted or engine_backward_and_forward" > $O/pytest_head_v12.log 2>&1; echo "pytest exit $?"; tail -6 $O/pytest_head_v12.log | cut -c1-400
timeout 200 python tools/head_bench.py > $O/head_bench.log 2>&1; timeout 200 python tools/head_bench.py --hidden 2048 --rows 32768 >> $O/head_bench.log 2>&1; cat $O/head_bench.log | cut -c1-900
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong"
for i in 1 2; do
DTA_FUSED_HEAD=0 python bench.py $F > $O/bench_ab4_unfused_$i.json 2> $O/bench_ab4_unfused_$i.err; echo "unfused exit $?"
DTA_FUSED_HEAD=1 python bench.py $F > $O/bench_ab4_fused_$i.json 2> $O/bench_ab4_fused_$i.err; echo "fused exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab4_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'], 'mem', round(d['peak_mem_gb'],1), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
