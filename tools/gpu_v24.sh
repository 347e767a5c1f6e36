#!/bin/bash
# Round-2 visit 24: the logits-keeping fused head launch: parity, then in-step A/B against the library GEMM + row kernel.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -k "fused or launches_counted or engine_backward_and_forward or 28_layer" > $O/pytest_head_v24.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_head_v24.log | cut -c1-300
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate"
i=0
for h in 0 auto 0 auto; do
  i=$((i+1))
  if [ $h = auto ]; then python bench.py $F > $O/bench_ab7_${i}_auto.json 2> /dev/null; else DTA_FUSED_HEAD=0 python bench.py $F > $O/bench_ab7_${i}_unfused.json 2> /dev/null; fi
  echo "$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab7_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'], 'mem', round(d['peak_mem_gb'],1), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
