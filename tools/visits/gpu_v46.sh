#!/bin/bash
# Round-2 visit 46: persistent forward with TWO softmax threads per row (DTA_FWD_SPLIT=1): parity, alone, in the step.
set -u
O=gpurun_out; mkdir -p $O
DTA_FWD_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "persistent or tree_attention_fwd_bwd or benchmarked or strided or chain_equals" > $O/pytest_split.log 2>&1
echo "pytest exit $?"; tail -12 $O/pytest_split.log
for e in 0 1; do
  DTA_FWD_SPLIT=$e timeout 300 python tools/attn_bench.py --case bushy --iters 7 > $O/attn_bushy_split$e.jsonl 2> $O/attn_bushy_split$e.err; echo "bushy split=$e exit $?"
done
python - <<'PY'
import json
a = [json.loads(l) for l in open('gpurun_out/attn_bushy_split0.jsonl')]
b = {d['case']: d for d in (json.loads(l) for l in open('gpurun_out/attn_bushy_split1.jsonl'))}
for d in a:
    e = b.get(d['case'])
    print(d['case'], 'fwd one-thread', d['fwd_persistent_tflops'], 'two-thread', e and e['fwd_persistent_tflops'])
PY
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate --no-vocab-parallel"
i=0
for h in 0 1 0 1; do
  i=$((i+1)); DTA_FWD_SPLIT=$h timeout 600 python bench.py $F > $O/bench_ab11_${i}_split${h}.json 2> /dev/null; echo "split=$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab11_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved'],1), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
