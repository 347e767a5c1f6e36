"""Debug: per-step timeline (cycles) of cluster 0 of the CTA-pair backward kernel (DTA_BWD_PAIR=1) on chain-16k, both CTAs,
next to the kernel times of the two variants on the same inputs.  clock64 is per SM: compare stamps within one CTA only.
Leader rows answer: how long do its issuers wait for the PEER after its own warps are done (C:P done -> p_full,
C:dS done -> ds_pair, D:ld done -> r1_free)?"""
import ctypes, os, statistics, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import torch
from dynamictreeattn_b200 import _cabi, attention
from dynamictreeattn_b200.attention import tree_attention
from dynamictreeattn_b200.linearize import linearize_trie
from dynamictreeattn_b200.token_trie import TokenTrie
from attn_bench import trunk_fan
lib = _cabi.load()
raw = ctypes.CDLL(str(_cabi.lib_path()))
trie = TokenTrie(trunk_fan(16384, 0, 1)); lay = linearize_trie(trie, "cuda")
N = lay.n_nodes
P = int(lay.pair_count.item())
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda h: torch.randn(N, h, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
q, k, v = mk(16), mk(8), mk(8); do = torch.randn(N, 16, 128, device="cuda", generator=g).bfloat16()
for mode in ("0", "1"):
    os.environ["DTA_BWD_PAIR"] = mode
    tree_attention(q, k, v, lay).backward(do)
    attention.PROFILE = []
    for _ in range(5):
        tree_attention(q, k, v, lay).backward(do)
    torch.cuda.synchronize()
    ms = statistics.median(a.elapsed_time(b) for kind, a, b in attention.PROFILE if kind == "bwd")
    attention.PROFILE = None
    print(f"chain16k backward main kernel, DTA_BWD_PAIR={mode}: {ms:.3f} ms, {10 * P * 16 * 128 / ms / 1e9:.1f} TFLOP/s")
os.environ["DTA_BWD_PAIR"] = "1"
buf = torch.zeros(48 * 32, dtype=torch.int64, device="cuda")
raw.dta_debug_set_bwd_trace(ctypes.c_void_p(buf.data_ptr()))
tree_attention(q, k, v, lay).backward(do)
torch.cuda.synchronize()
raw.dta_debug_set_bwd_trace(None)
os.environ["DTA_BWD_PAIR"] = "0"
t = buf.cpu().view(48, 2, 16)
names = ["p_full", "dV+St issued", "ds_full", "dQ+dK issued", "r1_free", "dP issued", "C:st_full", "C:P done", "C:dp_full",
         "C:dS done", "D:dq_full", "D:ld done", "D:red issued", "dV issued", "q_full seen", "ds_pair"]
for rank, who in ((0, "leader"), (1, "peer")):
    ref = 6 if rank == 1 else 0                       # the peer has no issuer A: anchor on its compute warp's st_full
    t0 = int(t[2, rank, ref])
    for it in range(2, 10):
        evs = range(16) if rank == 0 else (2, 6, 7, 8, 9, 10, 11, 12)
        print(f"{who} step {it}: " + "  ".join(f"{names[e]}={int(t[it, rank, e]) - t0}" for e in evs))
        t0 = int(t[it, rank, ref])
    print(f"{who} period:", [int(t[i + 1, rank, ref] - t[i, rank, ref]) for i in range(2, 24)])
