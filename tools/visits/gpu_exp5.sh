#!/bin/bash
set -u
O=gpurun_out/exp5.log; mkdir -p gpurun_out; : > $O
for f in 0 1 129 2 3 4 8 10 14 270 271 399; do
  echo "### flags $f" >> $O
  DTA_DEBUG_FLAGS=$f timeout 120 python tools/bwd_trace.py 2>&1 | tail -1 >> $O
done
cat $O
