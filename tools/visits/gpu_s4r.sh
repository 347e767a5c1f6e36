#!/bin/bash
# final-code records: config-5 sweep and ncu --set full of both attention kernels on the bench trie
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python tools/attn_bench.py --case sweep --iters 3 > $O/attn_sweep_final.jsonl 2>&1; tail -2 $O/attn_sweep_final.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tree_attn_.*_kernel -c 2 -f -o $O/attn_full_final \
    python tools/attn_bench.py --case tau2 --iters 1 > $O/ncu_full_final.log 2>&1; echo "ncu full exit $?"
ncu -i $O/attn_full_final.ncu-rep --page raw --csv > $O/attn_full_final_raw.csv 2>/dev/null; wc -c $O/attn_full_final_raw.csv
