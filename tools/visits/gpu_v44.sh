#!/bin/bash
# Round-2 visit 44: pair backward with the TMA-store dK / dV epilogue -- parity, alone, and in the step (alternating).
set -u
O=gpurun_out; mkdir -p $O
DTA_PAIR_TMA_EPILOGUE=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pair or benchmarked or qknorm_rope_attention" > $O/pytest_pair_tma.log 2>&1
echo "pytest exit $?"; tail -3 $O/pytest_pair_tma.log
for e in 0 1 0 1; do DTA_PAIR_TMA_EPILOGUE=$e timeout 200 python tools/attn_bench.py --case tau2 --iters 9 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('alone epilogue=$e bwd', d['bwd_tflops'], d['bwd_ms'], 'min', d['bwd_min_ms'])"; done
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate --no-vocab-parallel"
i=0
for h in 0 1 0 1; do
  i=$((i+1)); DTA_PAIR_TMA_EPILOGUE=$h timeout 600 python bench.py $F > $O/bench_ab10_${i}_tmaepi${h}.json 2> /dev/null; echo "tma_epilogue=$h exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab10_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],1), 'bwd', round(r['achieved'],1), 'fwd', round(r['attn_fwd_tflops']), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
