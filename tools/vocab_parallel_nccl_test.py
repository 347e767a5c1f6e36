"""Vocab-parallel logprob/entropy over NCCL (BASELINE config 4): run with torchrun --nproc-per-node {2,4,8}.

Each rank holds the [N, V/tp] shard of the same seeded logits; results and logits-gradients must match
the unsharded kernels and the fp32 oracle.  Also times the sharded op (HBM GB/s per rank).
"""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.distributed as dist
from dynamictreeattn_b200.vocab_parallel import gather_logprobs_entropy, gather_logprobs
from oracle import tree_oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V, N = 151936, int(os.environ.get("VP_ROWS", "4096"))
Vs = V // world
g = torch.Generator(device="cuda").manual_seed(0)
logits = (torch.randn(N, V, device="cuda", generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (N,), device="cuda", generator=g)
w1 = torch.randn(N, device="cuda", generator=g); w2 = torch.randn(N, device="cuda", generator=g)
shard = logits[:, rank * Vs:(rank + 1) * Vs].contiguous().requires_grad_(True)
lp, ent = gather_logprobs_entropy(shard, labels, temperature=0.9, tp_group=dist.group.WORLD)
((lp * w1).sum() + (ent * w2).sum()).backward()
full = logits.clone().requires_grad_(True)
lp1, ent1 = gather_logprobs_entropy(full, labels, temperature=0.9)          # tp=1 kernels
((lp1 * w1).sum() + (ent1 * w2).sum()).backward()
ref = logits.float().clone().requires_grad_(True)
lpr, entr = tree_oracle.logprob_entropy(ref, labels, 0.9)
((lpr * w1).sum() + (entr * w2).sum()).backward()
e_lp = max((lp - lp1).abs().max().item(), (lp - lpr).abs().max().item())
e_ent = max((ent - ent1).abs().max().item(), (ent - entr).abs().max().item())
gs = ref.grad[:, rank * Vs:(rank + 1) * Vs]
e_g = ((shard.grad.float() - gs).norm() / gs.norm()).item()
lp_only = gather_logprobs(shard.detach(), labels, temperature=0.9, tp_group=dist.group.WORLD)
e_lp2 = (lp_only - lpr).abs().max().item()
# timing
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    lp, ent = gather_logprobs_entropy(shard, labels, tp_group=dist.group.WORLD); (lp.sum() + ent.sum()).backward()
a.record()
iters = 10
for _ in range(iters):
    lp, ent = gather_logprobs_entropy(shard, labels, tp_group=dist.group.WORLD); (lp.sum() + ent.sum()).backward()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
t = torch.tensor([e_lp, e_ent, e_g, e_lp2, ms], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    e_lp, e_ent, e_g, e_lp2, ms = t.tolist()
    ok = e_lp < 1e-5 and e_ent < 5e-5 and e_g < 6e-3 and e_lp2 < 1e-5
    print(json.dumps({"test": "vocab_parallel_nccl", "tp": world, "rows": N, "ok": ok, "lp_err": e_lp, "ent_err": e_ent,
                      "grad_rel": e_g, "fwd_bwd_ms": ms, "gbs_per_rank": 6.0 * N * Vs / (ms * 1e-3) / 1e9}))
dist.destroy_process_group()
