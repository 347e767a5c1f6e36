#!/bin/bash
# session-4 visit B: backward timeline of HEAD, ncu --set full of both attention kernels (DRAM traffic), config-5 sweep
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python tools/bwd_trace.py > $O/bwd_trace_s4.log 2>&1; tail -12 $O/bwd_trace_s4.log
timeout 600 python tools/attn_bench.py --case sweep --iters 3 > $O/attn_sweep_s4.jsonl 2>&1; tail -5 $O/attn_sweep_s4.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tree_attn_.*_kernel -c 2 -f -o $O/attn_full_s4 \
    python tools/attn_bench.py --case tau2 --iters 1 > $O/ncu_full_s4.log 2>&1; echo "ncu full exit $?"
ncu -i $O/attn_full_s4.ncu-rep --page raw --csv > $O/attn_full_s4_raw.csv 2>/dev/null; wc -c $O/attn_full_s4_raw.csv
