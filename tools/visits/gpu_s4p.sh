#!/bin/bash
set -u
O=gpurun_out/s4p.log; mkdir -p gpurun_out; : > $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -2 >> $O
for c in "--case bary --b 64" "--case bary --b 8" "--case bary --b 2" "--case fan --p 1024 --b 64" "--case fan --p 4096 --b 16" "--case tau2"; do
  echo "### $c" >> $O; timeout 200 python tools/attn_bench.py $c --iters 6 2>&1 | tail -1 | cut -c1-260 >> $O; done
cat $O
