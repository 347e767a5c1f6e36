#!/bin/bash
set -u
O=gpurun_out/s4l.log; mkdir -p gpurun_out; : > $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $O
run() { echo "### $*" >> $O; env "$@" timeout 200 python tools/attn_bench.py --case tau2 --iters 8 2>&1 | tail -1 | cut -c60-260 >> $O; }
for rep in 1 2 3; do run DTA_BWD_SCHEDULE=whole; run DTA_BWD_SCHEDULE=worklist; done
echo "### bary64 (auto)" >> $O; timeout 200 python tools/attn_bench.py --case bary --b 64 --iters 6 2>&1 | tail -1 | cut -c1-260 >> $O
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> $O
cat $O
