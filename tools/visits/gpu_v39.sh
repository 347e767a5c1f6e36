#!/bin/bash
# Round-2 visit 39: persistent forward with the TMA-store epilogue and the causal-only diagonal path.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "persistent or tree_attention_fwd_bwd or benchmarked or strided or chain_equals" > $O/pytest_persist.log 2>&1
echo "pytest exit $?"; tail -5 $O/pytest_persist.log
for c in "--case fan --p 0 --b 64" "--case fan --p 1024 --b 64"; do timeout 200 python tools/fwd_persist_trace.py $c; done > $O/fwd_persist_trace2.log 2>&1
head -12 $O/fwd_persist_trace2.log
timeout 600 python tools/attn_bench.py --case ${SWEEP:-bushy} --iters 7 > $O/attn_sweep_persist.jsonl 2> $O/attn_sweep_persist.err
echo "sweep exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/attn_sweep_persist.jsonl'):
    d = json.loads(l)
    print(d['case'], 'auto', d['fwd_tflops'], 'one-item', d['fwd_one_item_per_cta_tflops'], 'persistent', d['fwd_persistent_tflops'], 'bwd', d['bwd_tflops'])
PY
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate --no-vocab-parallel"
i=0
for h in 0 1 0 1; do
  i=$((i+1)); DTA_FWD_PERSIST=$h timeout 600 python bench.py $F > $O/bench_ab9_${i}_persist${h}.json 2> /dev/null; echo "persist=$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab9_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved']), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
