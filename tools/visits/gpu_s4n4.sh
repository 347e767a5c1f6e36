#!/bin/bash
# 4-GPU visit: bench.py under torchrun (weak scaling + unequal forest) and the vocab-parallel NCCL test at config-4 size
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus 4 --steps 3 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 exit $?"
tail -1 $O/bench_n4.json | cut -c1-400; grep -v "OMP_NUM_THREADS\|^\*\*\*" $O/bench_n4.err | tail -3
VP_ROWS=16384 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 \
   tools/vocab_parallel_nccl_test.py > $O/vocab_parallel_tp4.log 2>&1; echo "vp exit $?"; tail -2 $O/vocab_parallel_tp4.log
