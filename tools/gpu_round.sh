#!/bin/bash
# One GPU-box visit: parity tests, bench, step breakdown, ncu launch list of the bench command, ncu --set full of the
# attention kernels.  Everything lands in gpurun_out/ (copied back by gpurun).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json
python tools/attn_bench.py --case tau2 > gpurun_out/attn_bench.json 2>&1; cat gpurun_out/attn_bench.json
python tools/step_profile.py > gpurun_out/step_profile.log 2>&1; head -40 gpurun_out/step_profile.log
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --no-cpu-baseline --no-dense-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tree_attn_.*_kernel -c 2 -f -o gpurun_out/attn_full \
    python tools/attn_bench.py --case tau2 --iters 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
fi
