"""Timeline of the persistent forward kernel's first items on CTA 0 (csrc/attn_fwd_persist.cu, kTrace instantiation).

    python tools/fwd_persist_trace.py [--case fan|bary|tau2] [--p 0] [--b 64]

Prints, per item, the clock64 stamps (cycles since the CTA's first stamp) of the producer (slot free, list published,
Q buffer free, loads issued, next item claimed), the MMA warp (item seen, Q landed, K landed, first P seen, last PV
issued) and one softmax warp per tile (item seen, first S seen, last P handed over, O final, epilogue done).
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from dynamictreeattn_b200.workloads import bary, synth_trie, trunk_fan  # noqa: E402

NAMES = {0: "slot_free", 1: "published", 2: "q_buf_free", 3: "loads_issued", 4: "next_claimed", 5: "mma_item", 20: "mma_q",
         6: "mma_k", 7: "mma_p0", 8: "mma_last_pv", 9: "A_item", 10: "A_s0", 11: "A_p_last", 12: "A_o_final", 13: "A_done",
         14: "B_item", 15: "B_s0", 16: "B_p_last", 17: "B_o_final", 18: "B_done"}
ORDER = [0, 1, 2, 3, 4, 5, 20, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="fan")
    ap.add_argument("--p", type=int, default=0)
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--n", type=int, default=16384)
    args = ap.parse_args()
    trace = torch.zeros(24 * 24, dtype=torch.int64, device="cuda")
    os.environ["DTA_FWD_PERSIST_TRACE"] = str(trace.data_ptr())
    os.environ["DTA_FWD_PERSIST"] = "1"
    from dynamictreeattn_b200 import attention
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.token_trie import TokenTrie
    seqs = {"fan": lambda: trunk_fan(args.n, args.p, args.b), "bary": lambda: bary(args.n, args.b, 1),
            "tau2": lambda: synth_trie(0, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384)}[args.case]()
    trie = TokenTrie(seqs)
    trie.backward_permute()
    lay = linearize_trie(trie, "cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(lay.n_nodes, 16, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(lay.n_nodes, 8, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(lay.n_nodes, 8, 128, device="cuda", generator=g).bfloat16()
    for _ in range(3):
        trace.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        attention.tree_attention_forward(q, k, v, lay, 128 ** -0.5)
        b.record()
        torch.cuda.synchronize()
    t = trace.cpu().view(24, 24)
    t0 = int(t[0, 0])
    print(f"case {args.case} p={args.p} b={args.b}: launch {a.elapsed_time(b) * 1e3:.1f} us; cycles since CTA 0's first stamp")
    print("item n_live " + " ".join(f"{NAMES[e]:>12}" for e in ORDER))
    for i in range(24):
        if int(t[i, 0]) == 0:
            break
        print(f"{i:4d} {int(t[i, 19]):6d} " + " ".join(f"{(int(t[i, e]) - t0) if int(t[i, e]) else -1:12d}" for e in ORDER))


if __name__ == "__main__":
    main()
