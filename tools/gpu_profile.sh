#!/bin/bash
# Round-2 profile visit: config-5 sweep with tile density, ncu --set full of the attention kernels and of the fused
# lm_head kernel, ncu launch list of a short bench run.  Each ncu pass only after the same command exited 0 without it.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python tools/attn_bench.py --case sweep --iters 3 > $O/attn_sweep_config5_r02.jsonl 2>&1; echo "sweep exit $?"; tail -3 $O/attn_sweep_config5_r02.jsonl | cut -c1-300
timeout 120 python tools/attn_bench.py --case tau2 --iters 1 > $O/attn_bench_tau2.json 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tree_attn_.*_kernel -c 2 -f -o $O/attn_full_r02 \
    python tools/attn_bench.py --case tau2 --iters 1 > $O/ncu_attn_full_r02.log 2>&1; echo "ncu attn exit $?"
ncu -i $O/attn_full_r02.ncu-rep --page raw --csv > $O/attn_full_r02_raw.csv 2>/dev/null; wc -c $O/attn_full_r02_raw.csv
timeout 120 python tools/head_bench.py --rows 8192 --iters 1 > $O/head_bench_8192.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lm_head_kernel -c 2 -f -o $O/head_full_r02 \
    python tools/head_bench.py --rows 8192 --iters 1 > $O/ncu_head_full_r02.log 2>&1; echo "ncu head exit $?"
ncu -i $O/head_full_r02.ncu-rep --page raw --csv > $O/head_full_r02_raw.csv 2>/dev/null; wc -c $O/head_full_r02_raw.csv
F="--no-cpu-baseline --no-dense-baseline --no-gpu-reference --no-strong --no-calibrate --steps 1 --warmup 3"
python bench.py $F > $O/bench_short.json 2> $O/bench_short.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02.csv -c 12000 \
    python bench.py $F > $O/ncu_bench_r02.log 2>&1; echo "ncu list exit $?"
python tools/launch_summary.py $O/launches_r02.csv > $O/launch_summary_r02.txt 2>&1; head -25 $O/launch_summary_r02.txt
gzip -f $O/launches_r02.csv
