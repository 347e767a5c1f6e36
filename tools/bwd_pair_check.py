"""A/B of the experimental CTA-pair attention backward (DTA_BWD_PAIR=1, csrc/attn_bwd_pair.cu) against the default
single-CTA kernel on the same inputs: dq / dk / dv differences, then kernel times on the tau2 shape.

    python tools/bwd_pair_check.py              # every case in order, stopping at the first failure
    python tools/bwd_pair_check.py --case NAME  # one case
    python tools/bwd_pair_check.py --sweep tau2 # kernel times of the pair variants (poly share x drain-arrival lane)

Cases grow from one key-tile pair to the bench shape so that the first failing line localises the bug.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CASES = ["chain_256", "chain_200", "chain_384", "chain_2048", "fan_small", "deep_forks", "many_leaves", "tau2_mini", "tau2"]


def seqs_for(name):
    import torch
    from trie_cases import synth_trie
    g = torch.Generator().manual_seed(0)
    if name.startswith("chain_"):
        return [torch.randint(0, 1000, (int(name.split("_")[1]),), generator=g)]
    if name == "fan_small":
        return synth_trie(1, 100, (3, 2), (40, 90), vocab=1000)
    if name == "deep_forks":
        return synth_trie(2, 300, (2, 2, 2, 2), (100, 260), vocab=1000)
    if name == "many_leaves":
        return synth_trie(8, 64, (8, 8, 4), (3, 40), vocab=3000)
    if name == "tau2_mini":
        return synth_trie(6, 1536, (4, 2, 2), (256, 768), vocab=5000, max_depth=4096)
    if name == "tau2":
        return synth_trie(0, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384)
    raise KeyError(name)


def run_case(name, Hq=16, Hkv=8, iters=3):
    import torch
    from dynamictreeattn_b200 import attention
    from dynamictreeattn_b200.attention import tree_attention
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.token_trie import TokenTrie
    trie = TokenTrie(seqs_for(name))
    trie.backward_permute()
    lay = linearize_trie(trie, "cuda")
    N = lay.n_nodes
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    k = torch.randn(N, Hkv, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    v = torch.randn(N, Hkv, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    do = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16()
    P = int(lay.pair_count.item())
    res = {"case": name, "N": N, "tiles": lay.n_pad // 128}
    grads, times = {}, {}
    timed = name == "tau2"
    for mode in ("0", "1"):
        os.environ["DTA_BWD_PAIR"] = mode
        attention.PROFILE = []
        for _ in range(iters if timed else 1):
            q.grad = k.grad = v.grad = None
            tree_attention(q, k, v, lay).backward(do)
        torch.cuda.synchronize()
        bwd = [a.elapsed_time(b) for kind, a, b in attention.PROFILE if kind == "bwd"]
        attention.PROFILE = None
        grads[mode] = [t.grad.float().clone() for t in (q, k, v)]
        times[mode] = statistics.median(bwd)
    os.environ["DTA_BWD_PAIR"] = "0"
    ok = True
    for nm, a, b in zip(("dq", "dk", "dv"), grads["0"], grads["1"]):
        rel = float((a - b).norm() / a.norm().clamp_min(1e-30))
        finite = bool(torch.isfinite(b).all())
        res[nm + "_rel"] = float(f"{rel:.3e}")
        res[nm + "_maxabs"] = float(f"{float((a - b).abs().max()):.3e}")
        ok = ok and finite and rel < 2e-3
    if not ok:
        # where it differs: per 64-row block (key / query halves of a tile) and per head-dim half (the pair's N split)
        for nm, a, b in zip(("dq", "dk", "dv"), grads["0"], grads["1"]):
            blocks = []
            for r0 in range(0, min(N, 768), 64):
                for c0 in (0, 64):
                    x, y = a[r0:r0 + 64, :, c0:c0 + 64], b[r0:r0 + 64, :, c0:c0 + 64]
                    blocks.append(round(float((x - y).norm() / x.norm().clamp_min(1e-30)), 3))
            res[nm + "_blocks_rows64_x_dhalf"] = blocks
    res["bwd_ms_single"] = round(times["0"], 3)
    res["bwd_ms_pair"] = round(times["1"], 3)
    if timed:
        res["bwd_tflops_single"] = round(10 * P * Hq * 128 / times["0"] / 1e9, 1)
        res["bwd_tflops_pair"] = round(10 * P * Hq * 128 / times["1"] / 1e9, 1)
    res["ok"] = ok
    print(json.dumps(res), flush=True)
    return ok


def sweep(name="tau2", Hq=16, Hkv=8, iters=5):
    """Kernel time of the pair variants (DTA_BWD_POLY x DTA_PAIR_DRAIN_LANE) next to the single-CTA kernel, one trie."""
    import torch
    from dynamictreeattn_b200 import attention
    from dynamictreeattn_b200.attention import tree_attention
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.token_trie import TokenTrie
    trie = TokenTrie(seqs_for(name))
    trie.backward_permute()
    lay = linearize_trie(trie, "cuda")
    N, P = lay.n_nodes, int(lay.pair_count.item())
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda h: torch.randn(N, h, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)   # noqa: E731
    q, k, v = mk(Hq), mk(Hkv), mk(Hkv)
    do = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16()
    variants = [("0", "0", "0", "arrive")] + [("1", po, la, "arrive") for po in ("0", "2", "4") for la in ("0", "1")]
    if os.environ.get("PAIR_SWEEP_COMMIT") == "1":           # only once tools/pair_handoff_bench mode 3 has passed: may trap
        variants += [("1", "0", "0", "commit"), ("1", "2", "0", "commit")]
    for pair, poly, lane, handoff in variants:
        os.environ.update(DTA_BWD_PAIR=pair, DTA_BWD_POLY=poly, DTA_PAIR_DRAIN_LANE=lane, DTA_PAIR_HANDOFF=handoff)
        tree_attention(q, k, v, lay).backward(do)
        attention.PROFILE = []
        for _ in range(iters):
            q.grad = k.grad = v.grad = None
            tree_attention(q, k, v, lay).backward(do)
        torch.cuda.synchronize()
        ms = statistics.median(a.elapsed_time(b) for kind, a, b in attention.PROFILE if kind == "bwd")
        attention.PROFILE = None
        print(json.dumps({"case": name, "pair": int(pair), "poly_of_8": int(poly), "drain_lane": int(lane), "handoff": handoff,
                          "bwd_ms": round(ms, 3), "bwd_tflops": round(10 * P * Hq * 128 / ms / 1e9, 1)}), flush=True)
    for key in ("DTA_BWD_PAIR", "DTA_BWD_POLY", "DTA_PAIR_DRAIN_LANE", "DTA_PAIR_HANDOFF"):
        os.environ.pop(key, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--sweep", default=None, help="time the pair variants on this case (e.g. tau2, chain_16384) and exit")
    ap.add_argument("--stop-on-fail", type=int, default=1)
    args = ap.parse_args()
    if args.sweep:
        sweep(args.sweep)
        return
    if args.case:
        sys.exit(0 if run_case(args.case) else 1)
    for name in CASES:       # in one process: after a trap the CUDA context is gone anyway, so stop at the first failure
        try:
            ok = run_case(name)
        except Exception as e:  # noqa: BLE001
            print(f"case {name} raised {type(e).__name__}: {str(e)[:300]}", flush=True)
            sys.exit(1)
        if not ok and args.stop_on_fail:
            sys.exit(1)


if __name__ == "__main__":
    main()
