"""Tree-attention kernel micro-benchmark (BASELINE config 5 sweep + the tau2 shape).

    python tools/attn_bench.py [--case tau2|chain|fan|bary|sweep] [--iters 5] [--p TRUNK] [--b FAN]

Times dta_tree_attn_fwd / dta_tree_attn_bwd alone (CUDA events around the main kernels, recorded
inside the C library) on Hq 16 / Hkv 8 / D 128 bf16 and reports TFLOP/s counting ONLY unmasked
(q,k) pairs: fwd 4*P*Hq*D, bwd 10*P*Hq*D.  Also the command to put under ncu.
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

import torch  # noqa: E402


from dynamictreeattn_b200.workloads import bary, synth_trie, trunk_fan  # noqa: E402,F401


def run_case(name, seqs, iters, Hq=16, Hkv=8, permute="backward"):
    from dynamictreeattn_b200 import attention
    from dynamictreeattn_b200.attention import tree_attention
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.token_trie import TokenTrie
    trie = TokenTrie(seqs)
    if permute == "backward":
        trie.backward_permute()
    elif permute == "forward":
        trie.forward_permute()
    lay = linearize_trie(trie, "cuda")
    N = lay.n_nodes
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    k = torch.randn(N, Hkv, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    v = torch.randn(N, Hkv, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    do = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16()
    P = int(lay.pair_count.item())
    for _ in range(2):
        tree_attention(q, k, v, lay).backward(do)
    attention.PROFILE = []
    for _ in range(iters):
        q.grad = k.grad = v.grad = None
        tree_attention(q, k, v, lay).backward(do)
    torch.cuda.synchronize()
    fwd = [a.elapsed_time(b) for kind, a, b in attention.PROFILE if kind == "fwd"]
    bwd = [a.elapsed_time(b) for kind, a, b in attention.PROFILE if kind == "bwd"]
    # the two forward kernels side by side (the call above takes the default: persistent up to 2048 tiles)
    fwd_ab = {}
    for mode in ("0", "1"):
        os.environ["DTA_FWD_PERSIST"] = mode
        with torch.no_grad():
            for _ in range(2):
                attention.tree_attention_forward(q, k, v, lay, 128 ** -0.5)
            attention.PROFILE = []
            for _ in range(max(iters, 5)):
                attention.tree_attention_forward(q, k, v, lay, 128 ** -0.5)
        torch.cuda.synchronize()
        fwd_ab[mode] = statistics.median(a.elapsed_time(b) for _, a, b in attention.PROFILE)
    os.environ.pop("DTA_FWD_PERSIST")
    attention.PROFILE = None
    f, b = statistics.median(fwd), statistics.median(bwd)
    # tile density: visible pairs / (live 128 x 128 tile pairs x 128^2) -- how much of the tensor work the kernels issue
    # is unmasked (the TFLOP/s below count ONLY visible pairs); partial = live tiles that need the per-element mask
    live = int(lay.q_nlive.sum().item())
    n_t = lay.n_pad // 128
    kmin = lay.kmin_end.cpu()
    kmax = lay.kmax_end.cpu()
    full = 0
    for qt in range(n_t):
        q0, q1 = qt * 128, qt * 128 + 128
        full += int(((kmax[:qt] > q0) & (kmin[:qt] >= q1)).sum())
    res = {"case": name, "N": N, "leaves": lay.n_leaf, "pairs": P, "live_tiles": live,
           "tile_density": round(P / (live * 16384.0), 4), "partial_tile_share": round(1.0 - full / max(live, 1), 4),
           "fwd_ms": round(f, 3), "bwd_ms": round(b, 3),
           "fwd_min_ms": round(min(fwd), 3), "bwd_min_ms": round(min(bwd), 3),
           "fwd_tflops": round(4 * P * Hq * 128 / f / 1e9, 1), "bwd_tflops": round(10 * P * Hq * 128 / b / 1e9, 1),
           "fwd_bwd_tflops": round(14 * P * Hq * 128 / (f + b) / 1e9, 1),
           "fwd_one_item_per_cta_tflops": round(4 * P * Hq * 128 / fwd_ab["0"] / 1e9, 1),
           "fwd_persistent_tflops": round(4 * P * Hq * 128 / fwd_ab["1"] / 1e9, 1)}
    print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="tau2")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--p", type=int, default=4096)
    ap.add_argument("--b", type=int, default=8)
    ap.add_argument("--n", type=int, default=16384)
    args = ap.parse_args()
    if args.case == "tau2":
        run_case("tau2", synth_trie(0, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384), args.iters)
    elif args.case == "chain":
        run_case(f"chain{args.n}", trunk_fan(args.n, 0, 1), args.iters)
    elif args.case == "fan":
        run_case(f"fan_p{args.p}_b{args.b}", trunk_fan(args.n, args.p, args.b), args.iters)
    elif args.case == "bary":
        import math
        run_case(f"bary_b{args.b}", bary(args.n, args.b, max(1, math.ceil(math.log(64, args.b)))), args.iters)
    elif args.case == "bushy":           # the short-item end of the sweep plus the two dense anchors
        import math
        for p, b in ((0, 1), (0, 8), (0, 16), (0, 32), (0, 64), (1024, 16), (1024, 64), (4096, 64)):
            run_case(f"fan_p{p}_b{b}", trunk_fan(args.n, p, b), args.iters)
        for b in (2, 8, 64):
            run_case(f"bary_b{b}", bary(args.n, b, max(1, math.ceil(math.log(64, b)))), args.iters)
        run_case("tau2", synth_trie(0, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384), args.iters)
    elif args.case == "sweep":
        import math
        for p in (0, 1024, 4096, 8192, 12288):
            for b in (1, 2, 4, 8, 16, 32, 64):
                run_case(f"fan_p{p}_b{b}", trunk_fan(args.n, p, b), args.iters)
        for b in (2, 4, 8, 64):
            run_case(f"bary_b{b}", bary(args.n, b, max(1, math.ceil(math.log(64, b)))), args.iters)
        run_case("tau2", synth_trie(0, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384), args.iters)


if __name__ == "__main__":
    main()
