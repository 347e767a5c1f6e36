"""-m gpu, >= 2 GPUs: the tp_group branch of vocab_parallel over NCCL (BASELINE config 4) — 2 ranks spawned from
the test process.  Skipped on a single-GPU box (the CPU suite covers the same branch under gloo:
tests/test_vocab_parallel_gloo.py); `gpurun --gpus 2 -- python -m pytest tests -m gpu -k nccl` runs it."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import json
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, str(ROOT))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    out = bench.vocab_parallel_leg(dev, 6539.2, rows=4096, iters=3)
    if rank == 0:
        Path(out_dir, "vp.json").write_text(json.dumps(out))
    dist.destroy_process_group()


def test_vocab_parallel_over_nccl_two_ranks(tmp_path):
    import json
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    out = json.loads((tmp_path / "vp.json").read_text())
    print("vocab-parallel tp=2:", out)
    assert out["ok"], out
