"""Debug: per-step timeline (cycles) of CTA 0 of the forward kernel on chain-16k, for each DTA_FWD_FLAGS setting.
CTA 0 is the heaviest query tile (q_order[0]); stamps: softmax warps of tile A / B, the MMA issuer."""
import ctypes, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

NAMES = ["A:s seen", "A:s loaded", "B:s seen", "B:s loaded", "M:pA half", "M:pA full", "M:A issued", "-",
         "M:pB half", "M:pB full", "M:B issued", "-", "A:p arrive", "B:p arrive", "A:p arr w3", "B:p arr w3"]


def one(flags: int):
    import torch
    from dynamictreeattn_b200 import _cabi
    from dynamictreeattn_b200.attention import tree_attention
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.workloads import trunk_fan
    _cabi.load()
    raw = ctypes.CDLL(str(_cabi.lib_path()))
    lay = linearize_trie(TokenTrie(trunk_fan(16384, 0, 1)), "cuda")
    N = lay.n_nodes
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda h: torch.randn(N, h, 128, device="cuda", generator=g).bfloat16()
    q, k, v = mk(16), mk(8), mk(8)
    tree_attention(q, k, v, lay)
    buf = torch.zeros(48 * 16, dtype=torch.int64, device="cuda")
    raw.dta_debug_set_fwd_trace(ctypes.c_void_p(buf.data_ptr()))
    tree_attention(q, k, v, lay)
    torch.cuda.synchronize()
    raw.dta_debug_set_fwd_trace(None)
    t = buf.cpu().view(48, 16)
    print(f"### DTA_FWD_FLAGS={flags}")
    for it in range(20, 26):
        t0 = int(t[it, 0])
        print(f"step {it}: " + "  ".join(f"{NAMES[e]}={int(t[it, e]) - t0}" for e in range(16) if NAMES[e] != "-" and int(t[it, e])))
    print("period (A sees S):", [int(t[i + 1, 0] - t[i, 0]) for i in range(8, 40)])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(int(sys.argv[1]))
    else:
        for f in (1, 9):
            subprocess.run([sys.executable, __file__, str(f)], env=dict(os.environ, DTA_FWD_FLAGS=str(f)))
