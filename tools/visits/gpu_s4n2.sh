#!/bin/bash
# 2-GPU visit: bench.py under torchrun (weak scaling, cost-model partition, NCCL grad all-reduce) + vocab-parallel NCCL test
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 exit $?"
tail -1 $O/bench_n2.json | cut -c1-600; tail -3 $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   tools/vocab_parallel_nccl_test.py > $O/vocab_parallel_tp2.log 2>&1; echo "vp exit $?"; tail -5 $O/vocab_parallel_tp2.log
