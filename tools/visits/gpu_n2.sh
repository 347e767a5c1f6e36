#!/bin/bash
# 2-GPU visit: vocab-parallel over NCCL (pytest, tp = 2) and the N = 2 bench line exactly as the driver launches it.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_vocab_parallel_nccl.py -m gpu -q -s > $O/pytest_nccl.log 2>&1; echo "pytest nccl exit $?"; grep -E "vocab-parallel|passed|failed" $O/pytest_nccl.log | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 exit $?"
tail -2 $O/bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
c=d['config']
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e'])
print('unequal', c['unequal_forest'])
print('strong', c['strong_scaling'])
print('dp_parity', c['dp_gradient_parity'])
print('vocab_parallel', c['vocab_parallel'])
PY
