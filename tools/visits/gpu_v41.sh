#!/bin/bash
# Round-2 visit 41: ncu --set full of the attention kernels with the persistent forward, and the ncu launch list of a
# short bench run (each only after the same command exited 0 without ncu).
set -u
O=gpurun_out; mkdir -p $O
timeout 120 python tools/attn_bench.py --case tau2 --iters 1 > $O/attn_bench_tau2.json 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tree_attn_.*_kernel -c 2 -f -o $O/attn_full_r02b \
    python tools/attn_bench.py --case tau2 --iters 1 > $O/ncu_attn_full_r02b.log 2>&1; echo "ncu attn exit $?"
ncu -i $O/attn_full_r02b.ncu-rep --page raw --csv > $O/attn_full_r02b_raw.csv 2>/dev/null; wc -c $O/attn_full_r02b_raw.csv
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate --no-vocab-parallel --steps 1 --warmup 3"
python bench.py $F > $O/bench_short.json 2> $O/bench_short.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02b.csv -c 12000 \
    python bench.py $F > $O/ncu_bench_r02b.log 2>&1; echo "ncu list exit $?"
python tools/launch_summary.py $O/launches_r02b.csv > $O/launch_summary_r02b.txt 2>&1; head -25 $O/launch_summary_r02b.txt
gzip -f $O/launches_r02b.csv
rm -f $O/attn_full_r02b.ncu-rep
