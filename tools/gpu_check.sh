#!/bin/bash
# parity tests + bench on the GPU box (no ncu)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python tools/step_profile.py > gpurun_out/step_profile.log 2>&1; head -32 gpurun_out/step_profile.log
