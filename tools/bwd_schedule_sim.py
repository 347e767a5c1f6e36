"""List-scheduling simulation of the attention backward launch (CPU only; DESIGN.md 3.2 "Long poles").

    python tools/bwd_schedule_sim.py [--sms 148] [--kv-heads 8] [--group 2]

For the bench trie and the config-5 sweep shapes: one job per (key tile, KV head) of n_qt * group steps (+ a fixed
per-CTA overhead), dispatched in k_order sequence to the first free SM — what the hardware block scheduler does with
one resident CTA per SM.  Prints the schedule efficiency (ideal / makespan) without and with the work list's split
rule (chunks of tiles longer than half a balanced share, chunk-major order).
"""
from __future__ import annotations

import argparse
import heapq
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))

from oracle import trie_oracle  # noqa: E402
from dynamictreeattn_b200.token_trie import TokenTrie  # noqa: E402


def jobs_of(seqs, n_kv, group, sms, split, overhead):
    trie = TokenTrie(seqs)
    trie.backward_permute()
    lin = trie_oracle.linearize([t.tolist() for t in trie.inputs], list(trie.lcp_lens))
    tm = trie_oracle.tile_meta(lin["subtree_end"])
    n_tiles = len(tm["k_qlast"])
    n_qt = [int(tm["k_qlast"][t] - t + 1) for t in range(n_tiles)]
    limit = 1 << 30
    if split:
        share = -(-sum(n_qt) * n_kv // sms)
        if max(n_qt) > max(16, (3 * share + 3) // 4):
            limit = max(16, (share + 1) // 2)
    parts = {t: -(-n_qt[t] // limit) for t in range(n_tiles)}
    order = [int(t) for t in tm["k_order"]]
    items = []
    for c in range(max(parts.values())):                    # chunk-major over the split tiles ...
        for t in order:
            if parts[t] > 1 and c < parts[t]:
                items.append((c + 1) * n_qt[t] // parts[t] - c * n_qt[t] // parts[t])
    items += [n_qt[t] for t in order if parts[t] == 1]      # ... then the whole tiles
    return [w * group + overhead for w in items for _ in range(n_kv)], n_tiles


def pair_jobs_of(seqs, n_kv, group, overhead):
    """CTA-pair kernel (csrc/attn_bwd_pair.cu): one job per (key tiles 2j / 2j+1, KV head) walking the UNION of the two
    live query ranges, in the order the pairs first appear in k_order.  Returns (jobs, single-tile steps)."""
    trie = TokenTrie(seqs)
    trie.backward_permute()
    lin = trie_oracle.linearize([t.tolist() for t in trie.inputs], list(trie.lcp_lens))
    tm = trie_oracle.tile_meta(lin["subtree_end"])
    n_tiles = len(tm["k_qlast"])
    qlast = [int(x) for x in tm["k_qlast"]]
    seen, items = set(), []
    for t in (int(t) for t in tm["k_order"]):
        j = t >> 1
        if j in seen:
            continue
        seen.add(j)
        last = max(qlast[2 * j], qlast[2 * j + 1] if 2 * j + 1 < n_tiles else 0)
        items.append(last - 2 * j + 1)
    single_steps = sum(qlast[t] - t + 1 for t in range(n_tiles)) * group * n_kv
    return [w * group + overhead for w in items for _ in range(n_kv)], single_steps


def efficiency(jobs, sms):
    heap = [0.0] * sms
    for j in jobs:
        heapq.heappush(heap, heapq.heappop(heap) + j)
    return sum(jobs) / sms / max(heap)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--kv-heads", type=int, default=8)
    ap.add_argument("--group", type=int, default=2)
    ap.add_argument("--overhead", type=float, default=1.5, help="per-CTA prologue + epilogue, in steps")
    a = ap.parse_args()
    from trie_cases import synth_trie
    from attn_bench import bary, trunk_fan
    cases = {"tau2 (bench)": synth_trie(0, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384),
             "chain 16k": trunk_fan(16384, 0, 1)}
    for b in (2, 8, 64):
        cases[f"bary b={b}"] = bary(16384, b, max(1, math.ceil(math.log(64, b))))
    for p, b in ((1024, 64), (4096, 16), (0, 64)):
        cases[f"fan p={p} b={b}"] = trunk_fan(16384, p, b)
    print(f"{'trie':<18s}{'tiles':>6s}{'jobs':>7s}{'eff (whole tiles)':>20s}{'jobs':>7s}{'eff (work list)':>18s}"
          f"{'pair: steps x':>15s}{'pair: eff':>11s}")
    for name, seqs in cases.items():
        j0, n_tiles = jobs_of(seqs, a.kv_heads, a.group, a.sms, False, a.overhead)
        j1, _ = jobs_of(seqs, a.kv_heads, a.group, a.sms, True, a.overhead)
        jp, single_steps = pair_jobs_of(seqs, a.kv_heads, a.group, a.overhead)
        # a pair step does the work of two single-CTA steps on two SMs: `steps x` = SM-steps spent / SM-steps needed
        waste = 2 * sum(j - a.overhead for j in jp) / single_steps
        print(f"{name:<18s}{n_tiles:>6d}{len(j0):>7d}{efficiency(j0, a.sms):>20.3f}{len(j1):>7d}{efficiency(j1, a.sms):>18.3f}"
              f"{waste:>15.3f}{efficiency(jp, a.sms // 2):>11.3f}")


if __name__ == "__main__":
    main()
