"""Vocab-parallel logprob/entropy over NCCL (BASELINE config 4): `torchrun --nproc-per-node {2,4,8} tools/vocab_parallel_nccl_test.py`
== `bench.py --vocab-parallel-only` (the leg also runs inside every multi-GPU `bench.py --gpus N` line as
`config.vocab_parallel`)."""
import runpy, sys
from pathlib import Path
sys.argv = [str(Path(__file__).resolve().parent.parent / "bench.py"), "--vocab-parallel-only"]
runpy.run_path(sys.argv[0], run_name="__main__")
