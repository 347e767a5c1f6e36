"""HBM GB/s of the log-prob / entropy row kernels alone (bench.py's `roofline.logprob` measurement), for A/B knobs:
    DTA_LOGPROB_THREADS=256|512|1024 python tools/logprob_bench.py [--rows 8192] [--vocab 151936]"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--vocab", type=int, default=151936)
a = ap.parse_args()
out = bench.logprob_roofline(torch.device("cuda", 0), bench.load_peaks()["hbm_gbs"], rows=a.rows, vocab=a.vocab, iters=7)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items() if k != "note"}))
