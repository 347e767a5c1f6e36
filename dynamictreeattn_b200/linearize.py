"""Device-side trie linearisation: (lens, lcp_lens) -> parent / depth / ancestor intervals.

The reference never materialises these arrays: its engine walks the trie with a
push/pop KV stack (tree_training_engine.py:541-546, 588-594).  The push order of that
walk IS the linear order used here, position id = depth, and the stack contents at
any time are exactly the ancestors of the current token, which gives the interval
mask ``k <= q < subtree_end[k]``.  All arrays are produced by the
``dta_trie_*`` kernels (csrc/trie_kernels.cu); nothing here computes them on the host.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi

TILE = _cabi.TILE


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class TrieLayout:
    """Linearised trie resident in HBM (all int32 unless noted)."""
    n_nodes: int
    n_pad: int
    n_leaf: int
    lens: List[int]
    lcp_lens: List[int]
    tok: torch.Tensor            # int64 [n_nodes]   token id of every node
    depth: torch.Tensor          # [n_pad]           position id
    parent: torch.Tensor         # [n_pad]           -1 at roots
    subtree_end: torch.Tensor    # [n_pad]           0 in the padding
    leaf_of: torch.Tensor        # [n_pad]
    leaf_off: torch.Tensor       # [n_leaf+1]        first node pushed by each leaf
    path: torch.Tensor           # [sum(lens)]       node at depth d of leaf i at path_off[i]+d
    path_off: List[int]
    kmax_end: torch.Tensor
    kmin_end: torch.Tensor
    q_order: torch.Tensor
    q_nlive: torch.Tensor
    k_order: torch.Tensor
    k_qlast: torch.Tensor
    pair_count: torch.Tensor     # int64 [1]  visible (q,k) pairs = sum(depth)+N
    _keep: list = field(default_factory=list, repr=False)
    _n_pairs: Optional[int] = field(default=None, repr=False)

    @property
    def n_tiles(self) -> int:
        return self.n_pad // TILE

    @property
    def n_pairs(self) -> int:
        """Visible (query, key) pairs = sum over nodes of depth + 1, from the host-side plan (no device read):
        leaf i pushes the depths lcp[i-1] .. lens[i]-1."""
        if self._n_pairs is None:
            from .trie import tree_pair_count
            self._n_pairs = int(tree_pair_count(self.lens, self.lcp_lens))
        return self._n_pairs

    def bwd_whole_tiles(self, n_kv_heads: int, n_sm: int) -> bool:
        """True when one CTA per (key tile, KV head) is provably balanced enough for the attention backward (see
        ``dta_attn_shape_t::bwd_whole_tiles``): no key tile's query range (at most n_tiles steps per head) exceeds
        half of an SM's share of the launch, bounded from below by pairs / 128^2 tile steps per KV head."""
        forced = os.environ.get("DTA_BWD_SCHEDULE")          # "whole" | "worklist": tests and A/B timing
        if forced in ("whole", "worklist"):
            return forced == "whole"
        share = (self.n_pairs // (TILE * TILE)) * n_kv_heads / max(n_sm, 1)
        return self.n_tiles <= 0.5 * share

    def meta(self) -> _cabi.AttnMeta:
        m = _cabi.AttnMeta()
        m.subtree_end = self.subtree_end.data_ptr()
        m.kmax_end = self.kmax_end.data_ptr()
        m.kmin_end = self.kmin_end.data_ptr()
        m.q_order = self.q_order.data_ptr()
        m.k_order = self.k_order.data_ptr()
        m.k_qlast = self.k_qlast.data_ptr()
        return m

    def path_of_leaf(self, i: int, length: int) -> torch.Tensor:
        """Node ids of the first ``length`` tokens on leaf i's root-to-leaf path."""
        o = self.path_off[i]
        return self.path[o:o + length]


def linearize(inputs: Sequence[torch.Tensor], lens: Sequence[int], lcp_lens: Sequence[int],
              device) -> TrieLayout:
    """Build the HBM layout for DFS-ordered leaves ``inputs`` (as exposed by TokenTrie)."""
    lib = _cabi.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise _cabi.DtaError("linearize() runs on the GPU only (no CPU fallback)")
    n_leaf = len(lens)
    if n_leaf == 0:
        raise ValueError("empty trie")
    lens = [int(x) for x in lens]
    lcp_lens = [int(x) for x in lcp_lens]
    starts = [0] + lcp_lens
    n_nodes = sum(lens) - sum(lcp_lens)
    n_pad = _round_up(n_nodes, TILE)
    total_path = sum(lens)

    # host staging: the new tokens of each leaf, back to back, plus the integer plan
    on_device = all(ids.is_cuda for ids in inputs)
    if on_device:           # CUDA-resident rollouts: the new tokens of each leaf are gathered on the device
        tok_host = None
    else:
        tok_host = torch.cat([ids[s:].reshape(-1).to("cpu", torch.int64) for ids, s in zip(inputs, starts)])
        assert tok_host.numel() == n_nodes
    plan = torch.tensor(lens + lcp_lens, dtype=torch.int32)
    path_off = [0] * (n_leaf + 1)
    for i, l in enumerate(lens):
        path_off[i + 1] = path_off[i] + l
    poff_host = torch.tensor(path_off, dtype=torch.int64)
    if torch.cuda.is_available():
        plan, poff_host = plan.pin_memory(), poff_host.pin_memory()
        if tok_host is not None:
            tok_host = tok_host.pin_memory()

    with torch.cuda.device(device):
        if on_device:
            tok = torch.cat([ids[s:].reshape(-1).to(device=device, dtype=torch.int64) for ids, s in zip(inputs, starts)])
            assert tok.numel() == n_nodes
        else:
            tok = tok_host.to(device, non_blocking=True)
        plan_d = plan.to(device, non_blocking=True)
        poff_d = poff_host.to(device, non_blocking=True)
        lens_d, lcp_d = plan_d[:n_leaf], plan_d[n_leaf:]
        i32 = dict(dtype=torch.int32, device=device)
        node_arrays = torch.empty((4, n_pad), **i32)
        depth, parent, subtree_end, leaf_of = node_arrays.unbind(0)
        leaf_off = torch.empty(n_leaf + 1, **i32)
        path = torch.empty(max(total_path, 1), **i32)
        ws = torch.empty(lib.dta_trie_workspace_bytes(n_leaf), dtype=torch.uint8, device=device)
        n_tiles = n_pad // TILE
        tiles = torch.empty((6, n_tiles), **i32)
        kmax_end, kmin_end, q_order, q_nlive, k_order, k_qlast = tiles.unbind(0)
        pair_count = torch.empty(1, dtype=torch.int64, device=device)
        st = _cabi.stream_ptr(device)
        _cabi.check(lib.dta_trie_linearize(lens_d.data_ptr(), lcp_d.data_ptr() if n_leaf > 1 else None, n_leaf,
                                           n_nodes, n_pad, leaf_off.data_ptr(), depth.data_ptr(),
                                           parent.data_ptr(), subtree_end.data_ptr(), leaf_of.data_ptr(),
                                           ws.data_ptr(), st))
        _cabi.check(lib.dta_trie_paths(lens_d.data_ptr(), lcp_d.data_ptr() if n_leaf > 1 else None,
                                       leaf_off.data_ptr(), n_leaf, poff_d.data_ptr(), total_path,
                                       path.data_ptr(), ws.data_ptr(), st))
        _cabi.check(lib.dta_tile_meta(subtree_end.data_ptr(), depth.data_ptr(), n_nodes, n_pad,
                                      kmax_end.data_ptr(), kmin_end.data_ptr(), q_order.data_ptr(),
                                      q_nlive.data_ptr(), k_order.data_ptr(), k_qlast.data_ptr(),
                                      pair_count.data_ptr(), st))
    return TrieLayout(n_nodes=n_nodes, n_pad=n_pad, n_leaf=n_leaf, lens=lens, lcp_lens=lcp_lens, tok=tok,
                      depth=depth, parent=parent, subtree_end=subtree_end, leaf_of=leaf_of, leaf_off=leaf_off,
                      path=path[:total_path], path_off=path_off, kmax_end=kmax_end, kmin_end=kmin_end,
                      q_order=q_order, q_nlive=q_nlive, k_order=k_order, k_qlast=k_qlast, pair_count=pair_count,
                      _keep=[plan_d, poff_d, ws, node_arrays, tiles])


def linearize_trie(token_trie, device) -> TrieLayout:
    return linearize(token_trie.inputs, token_trie.lens, token_trie.lcp_lens, device)


def lcp_adjacent_device(inputs: Sequence[torch.Tensor]) -> Tuple[List[int], bool]:
    """Adjacent LCPs + 'is some pair out of order' for CUDA-resident sequences (one launch)."""
    lib = _cabi.load()
    dev = inputs[0].device
    n = len(inputs)
    lens = [int(t.numel()) for t in inputs]
    off = [0] * (n + 1)
    for i, l in enumerate(lens):
        off[i + 1] = off[i] + l
    with torch.cuda.device(dev):
        tokens = torch.cat([t.reshape(-1).to(torch.int64) for t in inputs]) if off[-1] > 0 else \
            torch.zeros(1, dtype=torch.int64, device=dev)
        off_d = torch.tensor(off, dtype=torch.int64, device=dev)
        out = torch.empty(max(n - 1, 1) + 1, dtype=torch.int32, device=dev)
        _cabi.check(lib.dta_lcp_adjacent(tokens.data_ptr(), off_d.data_ptr(), n, out.data_ptr(),
                                         out[-1:].data_ptr(), _cabi.stream_ptr(dev)))
        host = out.cpu().tolist()
    return host[:n - 1], bool(host[-1])


MAX_DEVICE_BUILD_SEQ = 2048          # DTA_TRIE_BUILD_MAX_SEQ: the all-pairs tables are n_seq^2


def trie_build_device(inputs: Sequence[torch.Tensor], presorted: bool = False) -> dict:
    """Stable sort + leafization + forward / backward leaf orders of CUDA-resident rollouts in three launches
    (csrc/trie_build.cu), one D2H copy of the plan.  Returns Python lists:
    ``perm`` (sorted position -> input index), ``absorbed``, ``leaf_pos``, ``leaf_lcp``, ``order_fwd`` / ``lcp_fwd``,
    ``order_bwd`` / ``lcp_bwd`` and ``unsorted`` (adjacent input pairs out of order).  Bit-exact with the host code."""
    lib = _cabi.load()
    dev = inputs[0].device
    n = len(inputs)
    if n > MAX_DEVICE_BUILD_SEQ:
        raise _cabi.DtaError(f"device trie build handles up to {MAX_DEVICE_BUILD_SEQ} sequences (got {n})")
    lens = [int(t.numel()) for t in inputs]
    off = [0] * (n + 1)
    for i, l in enumerate(lens):
        off[i + 1] = off[i] + l
    with torch.cuda.device(dev):
        tokens = torch.cat([t.reshape(-1).to(torch.int64) for t in inputs]) if off[-1] > 0 else \
            torch.zeros(1, dtype=torch.int64, device=dev)
        off_d = torch.tensor(off, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
        ws = torch.empty(lib.dta_trie_build_workspace_bytes(n), dtype=torch.uint8, device=dev)
        out = torch.empty(2 + 9 * n, dtype=torch.int32, device=dev)
        _cabi.check(lib.dta_trie_build(tokens.data_ptr(), off_d.data_ptr(), n, 1 if presorted else 0, ws.data_ptr(),
                                       out.data_ptr(), _cabi.stream_ptr(dev)))
        host = out.cpu().tolist()
    n_leaf, unsorted = host[0], host[1]
    seg = lambda k, m: host[2 + k * n:2 + k * n + m]        # noqa: E731
    return {"n_leaf": n_leaf, "unsorted": unsorted, "perm": seg(0, n), "slcp": seg(1, max(n - 1, 0)),
            "absorbed": seg(2, n), "leaf_pos": seg(3, n_leaf), "leaf_lcp": seg(4, max(n_leaf - 1, 0)),
            "order_fwd": seg(5, n_leaf), "lcp_fwd": seg(6, max(n_leaf - 1, 0)),
            "order_bwd": seg(7, n_leaf), "lcp_bwd": seg(8, max(n_leaf - 1, 0)), "lens": lens}
