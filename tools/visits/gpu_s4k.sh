#!/bin/bash
# session-4 record visit: full GPU suite, bench, config-5 sweep, ncu launch list of the bench command (first 7000 launches)
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?"; cut -c1-300 $O/bench_n1.json
timeout 600 python tools/attn_bench.py --case sweep --iters 3 > $O/attn_sweep_s4b.jsonl 2>&1; tail -2 $O/attn_sweep_s4b.jsonl
python tools/step_profile.py > $O/step_profile.log 2>&1; head -12 $O/step_profile.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $O/launches_s4.csv \
    python bench.py --no-cpu-baseline > $O/ncu_bench_s4.log 2>&1; echo "ncu list exit $?"
python tools/launch_summary.py $O/launches_s4.csv > $O/launch_summary_s4.txt 2>&1; head -24 $O/launch_summary_s4.txt
gzip -f $O/launches_s4.csv
