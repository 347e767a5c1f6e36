#!/bin/bash
# round-end style check: GPU suite, smoke(), bench (N = 1), step profile
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?"; cut -c1-200 $O/bench_n1.json
python tools/step_profile.py > $O/step_profile.log 2>&1; grep -E "total kernel|qknorm|rmsnorm|delta|convert" $O/step_profile.log | head -12
