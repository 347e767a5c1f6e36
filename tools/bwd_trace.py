"""Debug: print the per-step timeline (cycles) of CTA 0 of the backward kernel."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from dynamictreeattn_b200 import _cabi
from dynamictreeattn_b200.attention import tree_attention
from dynamictreeattn_b200.linearize import linearize_trie
from dynamictreeattn_b200.token_trie import TokenTrie
sys.path.insert(0, str(ROOT / "tools"))
from attn_bench import trunk_fan
lib = _cabi.load()
raw = ctypes.CDLL(str(_cabi.lib_path()))
trie = TokenTrie(trunk_fan(16384, 0, 1)); lay = linearize_trie(trie, "cuda")
N = lay.n_nodes
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda h: torch.randn(N, h, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
q, k, v = mk(16), mk(8), mk(8); do = torch.randn(N, 16, 128, device="cuda", generator=g).bfloat16()
tree_attention(q, k, v, lay).backward(do)
buf = torch.zeros(48 * 16, dtype=torch.int64, device="cuda")
raw.dta_debug_set_bwd_trace(ctypes.c_void_p(buf.data_ptr()))
tree_attention(q, k, v, lay).backward(do)
torch.cuda.synchronize()
raw.dta_debug_set_bwd_trace(None)
t = buf.cpu().view(48, 16)
names = ["p_full", "dV+St issued", "ds_full", "dQ+dK issued", "r1_free", "dP issued", "C:st_full", "C:P done", "C:dp_full", "C:dS done", "D:dq_full", "D:ld done", "D:red issued", "dV issued", "q_full seen"]
t0 = int(t[2, 0])
for it in range(2, 12):
    print(f"step {it}: " + "  ".join(f"{names[e]}={int(t[it, e]) - t0}" for e in range(15)))
    t0 = int(t[it, 0])
print("period (p_full to p_full):", [int(t[i + 1, 0] - t[i, 0]) for i in range(2, 20)])
