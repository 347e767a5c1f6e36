#!/bin/bash
# Multi-GPU visit: the bench line exactly as the driver launches it (torchrun, one rank per GPU), N from $1.
set -u
N=${1:-4}
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N exit $?"
tail -2 $O/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
c=d['config']
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e'])
print('unequal', c['unequal_forest'])
print('strong', json.dumps(c['strong_scaling']))
print('dp_parity', c['dp_gradient_parity'])
print('vocab_parallel', c['vocab_parallel'])
print('tm', c['time_model'])
PY
