"""Per-kernel time breakdown of one Qwen3-0.6B trie step (torch.profiler, CUDA activities)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from transformers import Qwen3ForCausalLM
import bench
from dynamictreeattn_b200.linearize import linearize_trie
from dynamictreeattn_b200.token_trie import TokenTrie
from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine

dev = torch.device("cuda")
cfg = bench.qwen3_0p6b_config()
torch.manual_seed(0)
model = Qwen3ForCausalLM(cfg).to(dev, torch.bfloat16); model.train()
engine = TreeTrainingEngine(cfg, dev, torch.bfloat16, 16384)
seqs = bench.tau2_trie(0)
trie = TokenTrie(seqs, [dict(bench.ATTACH) for _ in seqs]); trie.backward_permute()
layout = linearize_trie(trie, dev)
for _ in range(2):
    engine.backward_layout(model, trie, layout, bench.example_loss, 2048)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    engine.backward_layout(model, trie, layout, bench.example_loss, 2048)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
# keep only kernels (drop aten parent ops): kernel names have no "aten::" prefix
kern = [r for r in rows if not r[0].startswith("aten::") and not r[0].startswith("autograd::") and "Backward" not in r[0]
        and not r[0].startswith("torch") and not r[0].startswith("_TreeHead") and not r[0].startswith("TreeAttn")]
kern.sort(key=lambda r: -r[1])
total = sum(r[1] for r in kern)
print(f"total kernel time {total:.1f} ms over {sum(r[2] for r in kern)} launches")
for k, ms, n in kern[:45]:
    print(f"{ms:9.2f} ms {100*ms/total:5.1f}% x{n:5d}  {k[:110]}")
