"""Trie-aware K-way load balancing + the NCCL data-parallel step.

Part 1 mirrors the reference module ``data_parallel.py`` (data_parallel.py:8-107):
``LB_by_n_tokens``, ``LB_by_TM``, ``LB_by_DFS_and_TM`` (+ ``pred_time``,
``get_original_bins``, ``try_devide``) return K lists of ORIGINAL sequence
indices, bit-identical to the reference for the same inputs and time model.

Part 2 is new: the reference only writes the bins to disk and replays them
serially on one GPU (exp/exp_dp.py:45-49).  Here every rank runs its own bin
through ``TreeTrainingEngine.backward`` and parameter gradients are summed with
bucketed NCCL all-reduce over NVLink.  The loss is a sum over
sequences (tree_training_engine.py:396-398), so any partition yields gradients
that add up to the single-GPU result.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .token_trie import TokenTrie
from .trie import CompressedTrie, _get_range_subtrie, _get_stats, _get_subtrie

__all__ = [
    "LB_by_n_tokens", "LB_by_TM", "LB_by_DFS_and_TM", "pred_time", "get_original_bins",
    "try_devide", "partition", "allreduce_gradients", "dp_backward", "GradientReducer",
]


# --------------------------------------------------------------------------
# Part 1 — partitioners (reference semantics)
# --------------------------------------------------------------------------
def _argmin_first(values: Sequence[float]) -> int:
    best = 0
    for j in range(1, len(values)):
        if values[j] < values[best]:
            best = j
    return best


def LB_by_n_tokens(token_seqs, K: int) -> List[List[int]]:
    """Greedy longest-first packing on raw token counts (data_parallel.py:8-16)."""
    bins: List[List[int]] = [[] for _ in range(K)]
    load = [0] * K
    for i in sorted(range(len(token_seqs)), key=lambda i: -len(token_seqs[i])):
        j = _argmin_first(load)
        bins[j].append(i)
        load[j] += len(token_seqs[i])
    return bins


def pred_time(compressed_trie: CompressedTrie, time_model, mode: str,
              block_size: Optional[int] = None) -> float:
    """Predicted step time of a (sub-)trie in its planned leaf order (data_parallel.py:18-27)."""
    if mode == "forward":
        _, lens, lcp_lens = compressed_trie.get_order_forward()
    elif mode == "backward":
        _, lens, lcp_lens = compressed_trie.get_order_backward()
    else:
        raise ValueError(f"Unsupported mode: {mode}")
    return time_model.pred(_get_stats(lens, lcp_lens, mode, block_size))


def _native_range_stats(compressed_trie: CompressedTrie, start: int, end: int, mode: str,
                        block_size: Optional[int]) -> Optional[dict]:
    """``_get_stats`` of the sub-trie of leaves [start, end) in its planned order, from the C++ host helper of the
    kernel library (``dta_host_subtrie_stats``: the same integer algorithm as trie.py, ~100x faster); None when the
    library is not built (the Python path then does the work)."""
    import ctypes as C
    try:
        import numpy as np
        from . import _cabi
        lib = _cabi.load(build_if_missing=False)
    except Exception:  # noqa: BLE001  (no library: pure-Python planning still works)
        return None
    cache = getattr(compressed_trie, "_np_leaves", None)
    if cache is None:
        cache = (np.asarray(compressed_trie.leaf_lens, dtype=np.int32), np.asarray(compressed_trie.leaf_lcps, dtype=np.int32))
        compressed_trie._np_leaves = cache
    lens, lcps = cache
    n = end - start
    out = np.empty(5, dtype=np.int64)
    lens_p = lens[start:end].ctypes.data_as(C.c_void_p)
    lcps_p = lcps[start:end - 1].ctypes.data_as(C.c_void_p) if n > 1 else None
    rc = lib.dta_host_subtrie_stats(lens_p, lcps_p, n, 1 if mode == "backward" else 0,
                                    int(block_size) if block_size else 0, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        return None
    stats = {"n_leaf_sequences": int(out[0]), "n_tree_tokens": int(out[1]), "sum_prefix_len": int(out[2]),
             "sum_depth": int(out[3])}
    if mode == "backward":
        stats["n_f1_tokens"] = int(out[4])
    return stats


def pred_time_range(compressed_trie: CompressedTrie, start: int, end: int, time_model, mode: str,
                    block_size: Optional[int] = None) -> float:
    """``pred_time(_get_subtrie(trie, set(range(start, end))), ...)`` for a contiguous range of the trie's leaves."""
    if mode not in ("forward", "backward"):
        raise ValueError(f"Unsupported mode: {mode}")
    stats = _native_range_stats(compressed_trie, start, end, mode, block_size)
    if stats is None:
        return pred_time(_get_range_subtrie(compressed_trie, start, end), time_model, mode, block_size)
    return time_model.pred(stats)


def get_original_bins(token_trie: TokenTrie, leaf_bins: List[List[int]]) -> List[List[int]]:
    """Leaf indices -> original sequence indices via the attach lists (data_parallel.py:29-37)."""
    return [
        [attach["_sequence_batch_id"] for leaf in leaf_bin for attach, _ in token_trie.attach_lists[leaf]]
        for leaf_bin in leaf_bins
    ]


def LB_by_TM(token_seqs, time_model, args) -> List[List[int]]:
    """Greedy: next leaf (DFS order) goes to the bin with least predicted time (data_parallel.py:39-56)."""
    token_trie = TokenTrie(token_seqs)
    full = CompressedTrie(token_trie.lens, token_trie.lcp_lens)
    K = args.K
    leaf_bins: List[List[int]] = [[] for _ in range(K)]
    bin_time = [0.0] * K
    for leaf in range(len(token_trie.inputs)):
        j = _argmin_first(bin_time)
        leaf_bins[j].append(leaf)
        bin_time[j] = pred_time(_get_subtrie(full, leaf_bins[j]), time_model, args.mode, args.block_size)
    return get_original_bins(token_trie, leaf_bins)


def try_devide(compressed_trie, n_seqs, args, divL, divR, time_model, cost_limit: float, _cost_cache=None):
    """Cut DFS-ordered leaves into consecutive runs of predicted cost <= cost_limit.

    Returns the run start indices; more than K of them means the limit is
    infeasible.  ``divL``/``divR`` bracket each cut from earlier probes
    (data_parallel.py:58-79).  ``_cost_cache`` (optional dict): predicted cost of a leaf range, shared by the probes
    of one ``LB_by_DFS_and_TM`` call -- the outer bisection asks for the same (start, end) ranges again and again; a
    pure memo, so the cuts are the reference's bit for bit.
    """
    K = args.K
    cuts: List[int] = []
    start = 0
    while start < n_seqs:
        cuts.append(start)
        if len(cuts) > K:
            break
        lo = max(divL[len(cuts)] - 1, start)
        hi = divR[len(cuts)] - 1
        while lo < hi:
            mid = (lo + hi + 1) // 2
            cost = _cost_cache.get((start, mid)) if _cost_cache is not None else None
            if cost is None:
                # == pred_time(_get_subtrie(trie, set(range(start, mid + 1))), ...)
                cost = pred_time_range(compressed_trie, start, mid + 1, time_model, args.mode, args.block_size)
                if _cost_cache is not None:
                    _cost_cache[(start, mid)] = cost
            if cost <= cost_limit:
                lo = mid
            else:
                hi = mid - 1
        start = lo + 1
    return cuts


def LB_by_DFS_and_TM(token_seqs, time_model, args) -> List[List[int]]:
    """Contiguous DFS ranges with minimal predicted makespan (data_parallel.py:81-107).

    Outer bisection on the per-bin cost limit between total/K and total, to
    1e-4 of the total; each probe is ``try_devide``.
    """
    token_trie = TokenTrie(token_seqs)
    n_leaf = len(token_trie.inputs)
    full = CompressedTrie(token_trie.lens, token_trie.lcp_lens)

    K = args.K
    hi = float(pred_time(full, time_model, args.mode, args.block_size))
    lo = hi / K
    eps = hi * 1e-4

    divL = [0] * (K + 1)
    divR = [n_leaf] * (K + 1)
    cache: dict = {}
    while hi - lo > eps:
        mid = (lo + hi) / 2.0
        cuts = try_devide(full, n_leaf, args, divL, divR, time_model, mid, cache)
        if len(cuts) <= K:
            hi = mid
            divR[: len(cuts)] = cuts
        else:
            lo = mid + eps
            divL = cuts[: K + 1]

    leaf_bins = [list(range(divR[i], divR[i + 1])) for i in range(K)]
    return get_original_bins(token_trie, leaf_bins)


def partition(token_seqs, K: int, method: str = "LB_by_DFS_and_TM", time_model=None,
              mode: str = "backward", block_size: Optional[int] = None) -> List[List[int]]:
    """Convenience front-end: pick a partitioner by the reference's CLI names."""
    from .tree_time_model import TreeTimeModel
    tm = time_model if time_model is not None else TreeTimeModel()
    args = SimpleNamespace(K=K, mode=mode, block_size=block_size)
    if method == "LB_by_n_tokens":
        return LB_by_n_tokens(token_seqs, K)
    if method == "LB_by_TM":
        return LB_by_TM(token_seqs, tm, args)
    if method == "LB_by_DFS_and_TM":
        return LB_by_DFS_and_TM(token_seqs, tm, args)
    raise ValueError(f"Unsupported method: {method}")


# --------------------------------------------------------------------------
# Part 2 — data-parallel runtime (new; one process per GPU, NCCL over NVLink)
# --------------------------------------------------------------------------
class GradientReducer:
    """Sum this step's parameter gradients over ranks, overlapped with the backward pass.

    One preallocated flat buffer per gradient dtype holds the step's gradients; while a step is armed every
    ``param.grad`` is a VIEW into it, so autograd accumulates in place and nothing is concatenated or copied back.
    The buffer is cut into buckets in REVERSE parameter order (the order the backward produces gradients); a
    post-accumulate-grad hook counts each parameter's arrivals and, when the last parameter of a bucket has
    received the gradient of the step's last pass, starts that bucket's all-reduce asynchronously (NCCL's own
    stream, ordered after the producing kernels) so that the exchange runs under the rest of the backward.  NVSwitch
    gives every GPU full bandwidth to every peer: buckets are sized for launch latency and overlap, not link count.

    The engine ACCUMULATES into ``param.grad`` and never zeroes it (reference tree_training_engine.py:487, 616), so
    gradients a caller already holds are set aside during the step and the reduced step gradient is added to them
    afterwards -- only the step's own contribution crosses the wire.

    ``arm(n_passes)``: a trie over the memory budget runs as several sub-tries (TreeTrainingEngine._ranges), each
    with its own backward; a bucket is reduced after the LAST of them.  ``finish()`` flushes what never fired
    (parameters without a gradient on this rank: an empty bin contributes zeros) and waits for the exchange.
    """

    def __init__(self, model: torch.nn.Module, group=None, bucket_bytes: int = 128 << 20):
        self.group = group
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat = {}            # dtype -> flat buffer
        self.slot = {}            # param -> (dtype, offset, numel)
        self.buckets = []         # (dtype, lo, hi, [params])
        self.bucket_of = {}
        by_dtype = {}
        for p in reversed(self.params):
            by_dtype.setdefault(p.dtype, []).append(p)
        for dtype, plist in by_dtype.items():
            total = sum(p.numel() for p in plist)
            self.flat[dtype] = torch.zeros(total, dtype=dtype, device=plist[0].device)
            off, lo, members = 0, 0, []
            esize = self.flat[dtype].element_size()
            for p in plist:
                self.slot[p] = (dtype, off, p.numel())
                members.append(p)
                off += p.numel()
                if (off - lo) * esize >= bucket_bytes:
                    self._close_bucket(dtype, lo, off, members)
                    lo, members = off, []
            if members:
                self._close_bucket(dtype, lo, off, members)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._armed = False
        self._kept = {}
        self._arrivals = {}
        self._pending = []
        self._works = []
        self._launched = set()
        self.n_passes = 1
        self.overlapped_buckets = 0     # buckets whose all-reduce started from a hook in the last step (diagnostic)

    def _close_bucket(self, dtype, lo, hi, members):
        idx = len(self.buckets)
        self.buckets.append((dtype, lo, hi, list(members)))
        for p in members:
            self.bucket_of[p] = idx

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    # ---- one step -----------------------------------------------------------------------------------------
    def arm(self, n_passes: int = 1) -> None:
        """Point every ``param.grad`` at its (zeroed) slot of the flat buffer; earlier gradients are kept aside."""
        self.n_passes = max(1, int(n_passes))
        self._kept = {}
        for p in self.params:
            dtype, off, n = self.slot[p]
            if p.grad is not None:
                mine = p.grad.data_ptr() == self.flat[dtype].data_ptr() + off * self.flat[dtype].element_size()
                # a gradient still living in the flat buffer (last step's, never zeroed by the caller) moves out first
                self._kept[p] = p.grad.clone() if mine else p.grad
        for buf in self.flat.values():
            buf.zero_()
        for p in self.params:
            dtype, off, n = self.slot[p]
            p.grad = self.flat[dtype][off:off + n].view_as(p)
        self._arrivals = {p: 0 for p in self.params}
        self._pending = [len(b[3]) for b in self.buckets]
        self._works, self._launched = [], set()
        self.overlapped_buckets = 0
        self._armed = True

    def set_passes(self, n_passes: int) -> None:
        """Called by the engine once it knows how many sub-tries (backward passes) this step runs as."""
        self.n_passes = max(1, int(n_passes))

    def _launch(self, idx: int) -> None:
        dtype, lo, hi, _ = self.buckets[idx]
        self._launched.add(idx)
        self._works.append(dist.all_reduce(self.flat[dtype][lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                           async_op=True))

    def _on_grad(self, p: torch.Tensor) -> None:
        if not self._armed:
            return
        self._arrivals[p] += 1
        if self._arrivals[p] == self.n_passes:
            idx = self.bucket_of[p]
            self._pending[idx] -= 1
            if self._pending[idx] == 0:
                self._launch(idx)
                self.overlapped_buckets += 1

    def finish(self) -> None:
        """Reduce what has not been reduced yet, wait, and hand the gradients back (added to any kept ones)."""
        if not self._armed:
            return
        self._armed = False
        for idx in range(len(self.buckets)):
            if idx not in self._launched:
                self._launch(idx)
        for w in self._works:
            w.wait()
        self._works = []
        for p, old in self._kept.items():
            old.add_(p.grad)
            p.grad = old
        self._kept = {}


_REDUCERS = {}


def _reducer_for(model, group) -> GradientReducer:
    key = (id(model), id(group))
    r = _REDUCERS.get(key)
    if r is None or [id(p) for p in r.params] != [id(p) for p in model.parameters() if p.requires_grad]:
        if r is not None:
            r.close()
        r = _REDUCERS[key] = GradientReducer(model, group)
    return r


def allreduce_gradients(model: torch.nn.Module, group=None, bucket_bytes: int = 256 << 20) -> None:
    """Sum ``param.grad`` over ranks AFTER the backward (no overlap): bucketed all-reduce of whatever the gradients
    hold.  Kept for callers that drive ``engine.backward`` themselves; ``dp_backward`` overlaps the exchange with
    the backward through ``GradientReducer`` instead.  Parameters whose grad is None on this rank contribute zeros."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for p in model.parameters() if p.requires_grad]
    by_dtype = {}
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        by_dtype.setdefault(p.grad.dtype, []).append(p)
    for plist in by_dtype.values():
        bucket, size = [], 0
        for p in plist + [None]:
            if p is not None:
                bucket.append(p)
                size += p.grad.numel() * p.grad.element_size()
            if bucket and (p is None or size >= bucket_bytes):
                flat = torch.cat([q.grad.reshape(-1) for q in bucket])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                off = 0
                for q in bucket:
                    n = q.grad.numel()
                    q.grad.copy_(flat[off:off + n].view_as(q.grad))
                    off += n
                bucket, size = [], 0


def dp_backward(engine, model, token_seqs, attachs, loss_fn: Callable, *, bins=None,
                method: str = "LB_by_DFS_and_TM", time_model=None, block_size: Optional[int] = 2048,
                group=None, permute: str = "ours", trie=None, layout=None, overlap: bool = True) -> float:
    """One data-parallel training step over a global batch of sequences.

    Every rank computes the SAME partition (deterministic host code, no broadcast needed), runs its bin through
    ``engine.backward`` and the step's gradients are summed over ranks (``GradientReducer``: flat buffer,
    reverse-order buckets, all-reduce overlapped with the backward).  Returns the global loss (sum over ranks).
    ``trie`` / ``layout``: this rank's bin already built / already resident in HBM (skips the host trie build /
    the H2D copy and linearisation: bench.py's device-resident timing)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if trie is None and token_seqs is not None:
        if bins is None:
            bins = partition(token_seqs, world, method, time_model, "backward", block_size) if world > 1 \
                else [list(range(len(token_seqs)))]
        mine = bins[rank]
        if mine:
            trie = TokenTrie([token_seqs[i] for i in mine], [dict(attachs[i]) for i in mine])
            if permute == "ours":
                trie.backward_permute()
    reducer = _reducer_for(model, group) if world > 1 and overlap else None
    if reducer is not None:
        reducer.arm(1)
        engine.pass_count_listener = reducer.set_passes
    loss = 0.0
    try:
        if trie is not None and trie.inputs:
            if layout is not None:
                loss = engine.backward_layout(model, trie, layout, loss_fn, block_size)
            else:
                loss = engine.backward(model, trie, loss_fn, block_size=block_size)
    finally:
        engine.pass_count_listener = None
        if reducer is not None:
            reducer.finish()
    if world > 1:
        if reducer is None:
            allreduce_gradients(model, group)
        t = torch.tensor([loss], dtype=torch.float64, device=next(model.parameters()).device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        loss = float(t.item())
    return loss
