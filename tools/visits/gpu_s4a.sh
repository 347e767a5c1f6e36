#!/bin/bash
# session-4 visit A: parity tests + bench + attention microbench on HEAD (no ncu)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python tools/attn_bench.py --case tau2 > gpurun_out/attn_bench.json 2>&1; tail -2 gpurun_out/attn_bench.json
python tools/attn_bench.py --case chain > gpurun_out/attn_bench_chain.json 2>&1; tail -2 gpurun_out/attn_bench_chain.json
python tools/step_profile.py > gpurun_out/step_profile.log 2>&1; head -40 gpurun_out/step_profile.log
