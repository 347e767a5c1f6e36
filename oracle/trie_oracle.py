"""ORACLE (test infrastructure, never shipped, never imported by the product package).

Naive CPU restatement of the reference's integer algorithms for the hot path, written
from the definitions rather than from either implementation, so that it can arbitrate
between the reference and the B200 code:

  lcp / leafization        reference token_trie.py:6-10, 13-49
  DFS leaf orders          reference trie.py:135-148, 171-202, 225-233 (explicit token-level
                           trie, recursive, only meant for small inputs)
  cost features            reference trie.py:7-56 (by enumerating every trie node)
  linearisation            the push/pop stack walk of reference
                           tree_training_engine.py:541-546, 588-594
  ancestor mask            parent-pointer closure

Pinned against the reference itself by tests/golden/*.json (generated in the build
container by tests/golden/make_golden.py importing /root/reference) and, when
/root/reference is present, by live differential tests.  Only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

from math import ceil
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------
# token_trie.py restated
# ----------------------------------------------------------------------------------------
def lcp(a: Sequence[int], b: Sequence[int]) -> int:
    n = 0
    for x, y in zip(a, b):
        if x != y:
            break
        n += 1
    return n


def sort_and_leafize(seqs: List[List[int]]):
    """-> (leaves, members, lcps): ``members[j]`` lists (original index, length) of every
    sequence folded into leaf j; raises ValueError if given unsorted input with presorted=True."""
    order = sorted(range(len(seqs)), key=lambda i: seqs[i])          # stable, list comparison
    srt = [seqs[i] for i in order]
    lcps = [lcp(srt[i], srt[i + 1]) for i in range(len(srt) - 1)]
    leaves, members, leaf_lcps = [], [], []
    pending: List[Tuple[int, int]] = []
    for i, s in enumerate(srt):
        pending.append((order[i], len(s)))
        is_prefix_of_next = i + 1 < len(srt) and lcps[i] >= min(len(s), len(srt[i + 1]))
        if not is_prefix_of_next:
            leaves.append(s)
            members.append(pending)
            pending = []
            if i + 1 < len(srt):
                leaf_lcps.append(lcps[i])
    return leaves, members, leaf_lcps


def check_sorted(seqs: List[List[int]]) -> bool:
    for a, b in zip(seqs, seqs[1:]):
        l = lcp(a, b)
        if l < min(len(a), len(b)) and a[l] > b[l]:
            return False
    return True


# ----------------------------------------------------------------------------------------
# trie.py restated on an explicit token-level trie
# ----------------------------------------------------------------------------------------
class _Node:
    __slots__ = ("depth", "kids", "leaf", "tail")

    def __init__(self, depth):
        self.depth = depth
        self.kids: List["_Node"] = []      # insertion (= lexicographic) order
        self.leaf = -1
        self.tail = 0


def _compressed(lens: Sequence[int], lcps: Sequence[int]) -> _Node:
    """Compressed trie from (lens, lcps) by inserting leaves along the right spine."""
    root = _Node(0)
    spine = [root]
    for i, n in enumerate(lens):
        f = lcps[i - 1] if i else 0
        last = None
        while len(spine) > 1 and (spine[-1].depth > f or spine[-1].leaf >= 0):
            last = spine.pop()
        if spine[-1].depth < f:
            mid = _Node(f)
            # `last` hangs below the top of the spine; re-hang it under the new fork node
            spine[-1].kids.remove(last)
            mid.kids.append(last)
            spine[-1].kids.append(mid)
            spine.append(mid)
        leaf = _Node(n)
        leaf.leaf = i
        spine[-1].kids.append(leaf)
        spine.append(leaf)
    return root


def leaf_order(lens: Sequence[int], lcps: Sequence[int], mode: str):
    """(order, lens, lcps) of the reference's forward / backward DFS."""
    root = _compressed(lens, lcps)
    if mode == "forward":
        key = lambda c: c.tail
    elif mode == "backward":
        key = lambda c: (1 if c.kids else 0, c.tail)
    else:
        raise ValueError(mode)

    def tails(v: _Node):
        if v.leaf >= 0:
            v.tail = v.depth
            return
        for c in v.kids:
            tails(c)
        v.tail = sorted(v.kids, key=key)[0].tail

    order, olens, olcps = [], [], []
    state = {"lca": 0}

    def walk(v: _Node):
        if v.leaf >= 0:
            order.append(v.leaf)
            olens.append(v.depth)
            olcps.append(state["lca"])
            state["lca"] = v.depth
            return
        for c in sorted(v.kids, key=key):
            state["lca"] = min(state["lca"], v.depth)
            walk(c)

    tails(root)
    walk(root)
    olcps = olcps[1:]
    if mode == "backward":
        return order[::-1], olens[::-1], olcps[::-1]
    return order, olens, olcps


def stats(lens: Sequence[int], lcps: Sequence[int], mode: str, block_size: Optional[int] = None) -> Dict[str, int]:
    """Cost features by enumerating every trie node (depth d contributes d to sum_depth)."""
    n_tree = 0
    sum_depth = 0
    for i, n in enumerate(lens):
        s = lcps[i - 1] if i else 0
        for d in range(s, n):
            n_tree += 1
            sum_depth += d
    out = {"n_leaf_sequences": len(lens), "n_tree_tokens": n_tree, "sum_depth": sum_depth}
    if mode == "forward":
        out["sum_prefix_len"] = sum(lcps)
        return out
    # backward: replay the reference engine's pop schedule (tree_training_engine.py:489-513, 594-609)
    spl = 0
    f1 = 0
    for i, n in enumerate(lens):
        keep = lcps[i] if i < len(lcps) else 0
        pushed_from = lcps[i - 1] if i else 0
        pop_len = n - keep
        if block_size is None or pop_len <= block_size:
            spl += keep
            f1_end = keep
        else:
            nb = ceil(pop_len / block_size)
            blk = ceil(pop_len / nb)
            f1_end = n - blk
            for b in range(nb):
                spl += max(n - (b + 1) * blk, keep)
        f1 += max(f1_end - pushed_from, 0)
    out["sum_prefix_len"] = spl
    out["n_f1_tokens"] = f1
    return out


# ----------------------------------------------------------------------------------------
# linearisation by stack simulation
# ----------------------------------------------------------------------------------------
def linearize(leaves: List[Sequence[int]], lcps: Sequence[int]):
    """Walk the leaves like the reference engine: truncate the stack to the fork depth, push
    the rest.  Returns dict of numpy arrays: tok, depth, parent, subtree_end, leaf_of,
    leaf_off, paths (list of per-leaf node-id arrays)."""
    tok, depth, parent, leaf_of = [], [], [], []
    stack: List[int] = []                       # node ids currently on the stack, by depth
    leaf_off = []
    paths = []
    for i, ids in enumerate(leaves):
        s = lcps[i - 1] if i else 0
        del stack[s:]
        leaf_off.append(len(tok))
        for d in range(s, len(ids)):
            n = len(tok)
            tok.append(int(ids[d]))
            depth.append(d)
            parent.append(stack[-1] if stack else -1)
            leaf_of.append(i)
            stack.append(n)
        paths.append(np.array(stack, dtype=np.int64))
    leaf_off.append(len(tok))
    N = len(tok)
    # subtree_end[n]: first later node whose depth <= depth[n]
    end = [N] * N
    open_nodes: List[int] = []
    for n in range(N):
        while open_nodes and depth[open_nodes[-1]] >= depth[n]:
            end[open_nodes.pop()] = n
        open_nodes.append(n)
    return {
        "tok": np.array(tok, dtype=np.int64),
        "depth": np.array(depth, dtype=np.int64),
        "parent": np.array(parent, dtype=np.int64),
        "subtree_end": np.array(end, dtype=np.int64),
        "leaf_of": np.array(leaf_of, dtype=np.int64),
        "leaf_off": np.array(leaf_off, dtype=np.int64),
        "paths": paths,
    }


def ancestor_mask(parent: np.ndarray) -> np.ndarray:
    """mask[q, k] = k is q or an ancestor of q — from parent pointers only."""
    N = len(parent)
    m = np.zeros((N, N), dtype=bool)
    for q in range(N):
        k = q
        while k >= 0:
            m[q, k] = True
            k = int(parent[k])
    return m


def interval_mask(subtree_end: np.ndarray) -> np.ndarray:
    N = len(subtree_end)
    q = np.arange(N)[:, None]
    k = np.arange(N)[None, :]
    return (k <= q) & (q < subtree_end[None, :])


K_ORDER_BUCKET = 64     # include/dta_b200.h DTA_K_ORDER_BUCKET


def tile_meta(subtree_end: np.ndarray, tile: int = 128):
    """Per-tile summaries the kernels consume (see include/dta_b200.h dta_tile_meta)."""
    N = len(subtree_end)
    n_tiles = (N + tile - 1) // tile
    kmax = np.zeros(n_tiles, dtype=np.int64)
    kmin = np.zeros(n_tiles, dtype=np.int64)
    for t in range(n_tiles):
        seg = subtree_end[t * tile:min((t + 1) * tile, N)]
        kmax[t], kmin[t] = seg.max(), seg.min()
    q_nlive = np.array([sum(1 for j in range(i + 1) if kmax[j] > i * tile) for i in range(n_tiles)], dtype=np.int64)
    k_qlast = (kmax - 1) // tile
    k_work = k_qlast - np.arange(n_tiles)
    q_order = np.array(sorted(range(n_tiles), key=lambda i: (-q_nlive[i], i)), dtype=np.int64)
    k_order = np.array(sorted(range(n_tiles), key=lambda i: (-(k_work[i] // K_ORDER_BUCKET), i)), dtype=np.int64)
    return {"kmax_end": kmax, "kmin_end": kmin, "q_nlive": q_nlive, "k_qlast": k_qlast,
            "q_order": q_order, "k_order": k_order}
