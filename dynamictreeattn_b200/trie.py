"""Compressed-trie planning over (lens, lcp_lens) — host side, integer, bit-exact.

Drop-in for the reference module ``trie.py`` (reference: trie.py:7-242).  The
public names and results are the reference's; the implementation is not: the
trie is held in flat parallel arrays and every traversal is iterative (the
reference recurses once per fork level, trie.py:135-188, and dies at Python's
recursion limit on deep fork chains).

Public surface:
    _get_stats(lens, lcp_lens, mode, block_size)      reference trie.py:7-56
    CompressedTrie(lens, lcp_lens)                    reference trie.py:66-91
        .get_order_forward()  -> (order, lens, lcp)   reference trie.py:225-228
        .get_order_backward() -> (order, lens, lcp)   reference trie.py:230-233
        .get_order_random(seed) -> order              reference trie.py:235-238
        .get_lens(seq_set) -> (lens, lcp)             reference trie.py:164-169
    _get_subtrie(trie, seq_set)                       reference trie.py:240-242
"""
from __future__ import annotations

import random
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

__all__ = ["_get_stats", "CompressedTrie", "_get_subtrie", "_get_range_subtrie", "tree_pair_count"]


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def tree_pair_count(lens: Sequence[int], lcp_lens: Sequence[int]) -> int:
    """Number of (query, key) pairs the ancestor mask leaves visible.

    Equals ``sum_depth + n_tree_tokens`` of ``_get_stats``: a node
    at depth d attends to d + 1 keys.  This is the FLOP denominator of the
    attention roofline.
    """
    pairs = 0
    for i, end in enumerate(lens):
        start = lcp_lens[i - 1] if i > 0 else 0
        n = end - start
        pairs += (start + end - 1) * n // 2 + n
    return pairs


def _get_stats(lens: Sequence[int], lcp_lens: Sequence[int], mode: str,
               block_size: Optional[int] = None) -> dict:
    """Cost-model features of a DFS-ordered leaf list (reference trie.py:7-56).

    forward : n_leaf_sequences, n_tree_tokens, sum_prefix_len, sum_depth
    backward: the same plus n_f1_tokens; sum_prefix_len/n_f1_tokens follow the
              reference engine's block-wise pop schedule (trie.py:26-46).

    Deliberate difference: the reference reads ``lcp_lens[i]`` for the LAST leaf
    when its pop fits one block (trie.py:35) and raises IndexError; the value it
    means is 0 (tree_training_engine.py:597), which is what is used here.
    """
    n_leaf = len(lens)
    n_tree_tokens = sum(lens) - sum(lcp_lens)
    sum_depth = 0
    for i in range(n_leaf):
        lo = lcp_lens[i - 1] if i > 0 else 0
        hi = lens[i]
        sum_depth += (lo + hi - 1) * (hi - lo) // 2

    if mode == "forward":
        return {
            "n_leaf_sequences": n_leaf,
            "n_tree_tokens": n_tree_tokens,
            "sum_prefix_len": sum(lcp_lens),
            "sum_depth": sum_depth,
        }
    if mode != "backward":
        raise ValueError(f"Unsupported mode: {mode}")

    sum_prefix_len = 0
    n_f1_tokens = 0
    for i in range(n_leaf):
        keep = lcp_lens[i] if i < len(lcp_lens) else 0      # depth the stack is cut back to
        hi = lens[i]
        popped = hi - keep
        pushed_from = lcp_lens[i - 1] if i > 0 else 0
        if block_size is None or popped <= block_size:
            f1_end = keep
            sum_prefix_len += keep
        else:
            n_blocks = _ceil_div(popped, block_size)
            blk = _ceil_div(popped, n_blocks)
            f1_end = hi - blk
            for b in range(n_blocks):
                sum_prefix_len += max(hi - (b + 1) * blk, keep)
        n_f1_tokens += max(f1_end - pushed_from, 0)

    return {
        "n_leaf_sequences": n_leaf,
        "n_tree_tokens": n_tree_tokens,
        "sum_prefix_len": sum_prefix_len,
        "sum_depth": sum_depth,
        "n_f1_tokens": n_f1_tokens,
    }


class CompressedTrie:
    """Compressed trie of DFS-ordered leaves, used to plan the leaf visiting order.

    Node 0 is the root (depth 0).  Internal nodes sit at fork depths, leaves at
    their sequence length.  Storage is struct-of-arrays:
        depth[v]     node depth
        seq_id[v]    leaf index, -1 for internal nodes
        children[v]  child node ids in lexicographic (input) order
        tail[v]      "chain tail depth" = depth of the leaf reached by always
                     following the first child under the active child order
    """

    def __init__(self, lens: Sequence[int], lcp_lens: Sequence[int]):
        if len(lcp_lens) != len(lens) - 1:
            raise ValueError("lcp_lens的长度必须为len(lens)-1 (len(lcp_lens) must be len(lens)-1)")
        self.leaf_lens = list(lens)          # the DFS-ordered leaves this trie was built from
        self.leaf_lcps = list(lcp_lens)
        self.depth: List[int] = [0]
        self.seq_id: List[int] = [-1]
        self.children: List[List[int]] = [[]]
        self.tail: List[int] = [0]
        self._build(lens, lcp_lens)
        # results of the last traversal, kept under the reference's attribute names
        self.order = None
        self.lens = None
        self.lcp_lens = None

    # ---- construction -----------------------------------------------------
    def _add(self, depth: int, seq_id: int = -1) -> int:
        self.depth.append(depth)
        self.seq_id.append(seq_id)
        self.children.append([])
        self.tail.append(0)
        return len(self.depth) - 1

    def _build(self, lens: Sequence[int], lcp_lens: Sequence[int]) -> None:
        """Right-spine construction (same trie as reference trie.py:93-133).

        ``spine`` holds the root→current-leaf path.  For a new leaf with fork
        depth f: close every spine node deeper than f (attach it to the node
        below it); if no spine node sits exactly at depth f, splice an internal
        node there, adopting the last closed node; then open the new leaf.
        """
        depth, children, seq_id = self.depth, self.children, self.seq_id
        spine = [0]
        for i, length in enumerate(lens):
            f = lcp_lens[i - 1] if i > 0 else 0
            closed = -1
            # the previous leaf is always closed, even in the degenerate lens[i-1] == f case
            while len(spine) > 1 and (depth[spine[-1]] > f or seq_id[spine[-1]] != -1):
                v = spine.pop()
                if closed >= 0:
                    children[v].append(closed)
                closed = v
            if depth[spine[-1]] < f:
                spine.append(self._add(f))
            if closed >= 0:
                children[spine[-1]].append(closed)
            spine.append(self._add(length, i))
        while len(spine) > 1:
            v = spine.pop()
            children[spine[-1]].append(v)

    # ---- child orders ------------------------------------------------------
    def _order_forward(self, v: int) -> List[int]:
        tail = self.tail
        return sorted(self.children[v], key=lambda c: tail[c])

    def _order_backward(self, v: int) -> List[int]:
        tail, children = self.tail, self.children
        return sorted(self.children[v], key=lambda c: (1 if children[c] else 0, tail[c]))

    def _order_random(self, v: int, seed: Optional[int]) -> List[int]:
        kids = list(self.children[v])
        (random.Random(seed) if seed is not None else random).shuffle(kids)
        return kids

    # ---- traversals --------------------------------------------------------
    def _compute_tails(self, child_order: Callable[[int], List[int]]) -> None:
        """Post-order pass: tail[v] = tail of the first child under ``child_order``.

        The order function is invoked once per internal node, after all of its
        descendants (reference trie.py:135-148) — this matters for the unseeded
        random order, which consumes the global RNG stream.
        """
        seq_id, children, tail, depth = self.seq_id, self.children, self.tail, self.depth
        todo = [(0, 0)]
        while todo:
            v, k = todo.pop()
            if seq_id[v] != -1:
                tail[v] = depth[v]
                continue
            kids = children[v]
            if k < len(kids):
                todo.append((v, k + 1))
                todo.append((kids[k], 0))
            else:
                tail[v] = tail[child_order(v)[0]]

    def _walk(self, child_order: Callable[[int], List[int]], keep=None):
        """Pre-order pass emitting (order, lens, lcp) of the visited leaves.

        ``lcp`` of a leaf is the shallowest node depth crossed since the previous
        emitted leaf (reference trie.py:171-188 / 150-162).
        """
        seq_id, depth = self.seq_id, self.depth
        order: List[int] = []
        lens: List[int] = []
        lcps: List[int] = []
        lca = 0
        todo: List[Tuple[int, int]] = [(0, -1)]          # (node, parent depth or -1)
        while todo:
            v, pd = todo.pop()
            if pd >= 0 and pd < lca:
                lca = pd
            s = seq_id[v]
            if s != -1:
                if keep is None or s in keep:
                    order.append(s)
                    lens.append(depth[v])
                    lcps.append(lca)
                    lca = depth[v]
                continue
            kids = child_order(v)
            d = depth[v]
            for c in reversed(kids):
                todo.append((c, d))
        return order, lens, lcps

    def get_order(self, child_order: Callable[[int], List[int]]) -> None:
        self._compute_tails(child_order)
        self.order, self.lens, self.lcp_lens = self._walk(child_order)

    def get_order_forward(self):
        """Shortest-chain-first DFS (reference trie.py:225-228)."""
        self.get_order(self._order_forward)
        return self.order, self.lens, self.lcp_lens[1:]

    def get_order_backward(self):
        """Leaves-then-forks, shortest first, whole order reversed (trie.py:230-233)."""
        self.get_order(self._order_backward)
        return self.order[::-1], self.lens[::-1], self.lcp_lens[1:][::-1]

    def get_order_random(self, seed: Optional[int] = None):
        self.get_order(lambda v: self._order_random(v, seed))
        return self.order

    def get_lens(self, seq_set: Iterable[int]):
        """(lens, lcp_lens) of the sub-trie spanned by the leaves in ``seq_set``."""
        keep = seq_set if isinstance(seq_set, (set, frozenset)) else set(seq_set)
        _, self.lens, self.lcp_lens = self._walk(lambda v: self.children[v], keep)
        return self.lens, self.lcp_lens[1:]

    # The reference exposes ``nodes``; keep a light read-only equivalent.
    @property
    def nodes(self):
        return [
            {"depth": d, "seq_id": s, "chain_tail_depth": t, "child_ids": list(c)}
            for d, s, t, c in zip(self.depth, self.seq_id, self.tail, self.children)
        ]


def _get_subtrie(trie: CompressedTrie, seq_set) -> CompressedTrie:
    lens, lcp_lens = trie.get_lens(seq_set)
    return CompressedTrie(lens, lcp_lens)


def _get_range_subtrie(trie: CompressedTrie, start: int, end: int) -> CompressedTrie:
    """``_get_subtrie(trie, set(range(start, end)))`` for a CONTIGUOUS range of the trie's own (DFS-ordered) leaves:
    nothing is skipped between two kept leaves, so the sub-trie's lens / lcp_lens are slices of the trie's inputs and
    the O(n_leaf) walk of ``get_lens`` is not needed (same result, asserted by tests/test_host_golden.py)."""
    return CompressedTrie(trie.leaf_lens[start:end], trie.leaf_lcps[start:end - 1])
