"""TokenTrie — merge rollouts that share a prefix into DFS-ordered trie leaves.

Drop-in for the reference module ``token_trie.py`` (reference token_trie.py:1-109):
same constructor, attributes (``inputs attach_lists lcp_lens lens n_sequences
n_tokens``), methods and error contract.  Integer results are bit-exact with the
reference; the implementation differs:

* sort keys are big-endian byte strings compared with memcmp instead of Python
  lists (identical order for non-negative ids, including "shorter prefix first"
  and stability on duplicates; falls back to list keys if an id is negative);
* longest-common-prefix runs on numpy views in one vectorised compare per
  pair, no device sync per pair (reference: one ``.item()`` each,
  token_trie.py:6-10).  When the sequences already live on a CUDA device the
  adjacent-LCP + sortedness check runs in one launch of the ``dta_lcp_adjacent``
  kernel instead;
* rollouts that live on a CUDA device (the usual case in RL: they were just generated there) never come back to
  the host token by token: sort, leafization and BOTH leaf orders are computed by ``dta_trie_build``
  (csrc/trie_build.cu: all-pairs LCP + rank sort + one planning CTA) and one small plan is copied back
  (SURVEY.md section 8f, N2).  ``forward_permute`` / ``backward_permute`` then apply the precomputed order.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .trie import CompressedTrie, _get_stats

__all__ = ["TokenTrie", "_lcp_torch", "_leafization"]


def _as_np(ids: torch.Tensor) -> np.ndarray:
    t = ids.detach()
    if t.device.type != "cpu":
        t = t.cpu()
    return t.contiguous().numpy()


def _lcp_np(a: np.ndarray, b: np.ndarray) -> int:
    n = min(a.shape[0], b.shape[0])
    if n == 0:
        return 0
    diff = a[:n] != b[:n]
    hit = int(diff.argmax())
    return hit if diff[hit] else n


def _lcp_torch(a: torch.Tensor, b: torch.Tensor) -> int:
    """Length of the longest common prefix of two 1-D id tensors (token_trie.py:6-10)."""
    return _lcp_np(_as_np(a), _as_np(b))


def _adjacent_lcps(views: Sequence[np.ndarray], inputs=None, check_sorted: bool = True) -> List[int]:
    """LCP of every adjacent pair; raises ValueError if a pair is out of order."""
    if inputs is not None and len(inputs) > 1 and inputs[0].is_cuda:
        from . import linearize  # CUDA path; fails loudly if the extension is missing
        lcps, bad = linearize.lcp_adjacent_device(inputs)
        if check_sorted and bad:
            raise ValueError("Input_ids not sorted in lexicographic order.")
        return lcps
    out = []
    for i in range(len(views) - 1):
        a, b = views[i], views[i + 1]
        l = _lcp_np(a, b)
        if check_sorted and l < min(a.shape[0], b.shape[0]) and a[l] > b[l]:
            raise ValueError("Input_ids not sorted in lexicographic order.")
        out.append(l)
    return out


def _leafization(input_ids: List[torch.Tensor], attachs: List[dict]):
    """Collapse sequences that are full prefixes of their successor (token_trie.py:13-49).

    Input must be lexicographically sorted.  Returns ``(leaves, attach_lists,
    lcp_lens)`` where ``attach_lists[j]`` lists ``(attach, length)`` for every
    original sequence folded into leaf j (the leaf itself last).
    """
    views = [_as_np(t) for t in input_ids]
    n = len(input_ids)
    lcps = _adjacent_lcps(views, input_ids)

    leaves, attach_lists, leaf_lcps = [], [], []
    first = 0                                   # first sequence not yet folded into a leaf
    for i in range(n):
        absorbed = i + 1 < n and lcps[i] >= min(views[i].shape[0], views[i + 1].shape[0])
        if absorbed:
            continue
        leaves.append(input_ids[i])
        if i + 1 < n:
            leaf_lcps.append(lcps[i])
        attach_lists.append([(attachs[k], int(views[k].shape[0])) for k in range(first, i + 1)])
        first = i + 1
    return leaves, attach_lists, leaf_lcps


def _sort_keys(views: Sequence[np.ndarray]):
    if all(v.size == 0 or int(v.min()) >= 0 for v in views):
        return [v.astype(">i8", copy=False).tobytes() for v in views]
    return [v.tolist() for v in views]


class TokenTrie:
    def __init__(self, inputs: List[torch.Tensor], attachs: Optional[List[dict]] = None,
                 sorted: bool = False):
        if attachs is not None:
            assert len(inputs) == len(attachs), "Length of inputs and attachs must match."
        else:
            attachs = [{} for _ in range(len(inputs))]

        # the caller's dicts are tagged in place, as the reference does (token_trie.py:64-65)
        for seq_id in range(len(inputs)):
            attachs[seq_id]["_sequence_batch_id"] = seq_id

        self._device_plan = None
        if self._use_device_build(inputs):
            self._init_from_device(inputs, attachs, sorted)
            return

        if not sorted:
            keys = _sort_keys([_as_np(t) for t in inputs])
            perm = list(range(len(inputs)))
            perm.sort(key=keys.__getitem__)       # stable, like list.sort on (ids, attach) pairs
            inputs_sorted = [inputs[i] for i in perm]
            attachs_sorted = [attachs[i] for i in perm]
        else:
            inputs_sorted, attachs_sorted = inputs, attachs

        self.inputs, self.attach_lists, self.lcp_lens = _leafization(inputs_sorted, attachs_sorted)
        self.lens = [len(ids) for ids in self.inputs]

        self.n_sequences = len(inputs)
        self.n_tokens = sum(len(ids) for ids in inputs)

    # ---- device-side build (CUDA-resident rollouts) -------------------------
    @staticmethod
    def _use_device_build(inputs) -> bool:
        if len(inputs) < 2 or not all(torch.is_tensor(t) and t.is_cuda and t.dim() == 1 for t in inputs):
            return False
        from . import linearize
        return len(inputs) <= linearize.MAX_DEVICE_BUILD_SEQ

    def _init_from_device(self, inputs, attachs, presorted: bool) -> None:
        from . import linearize
        plan = linearize.trie_build_device(inputs, presorted)
        if presorted and plan["unsorted"]:
            raise ValueError("Input_ids not sorted in lexicographic order.")
        perm, lens = plan["perm"], plan["lens"]
        self.inputs, self.attach_lists = [], []
        members = []
        for r, src in enumerate(perm):                  # sorted order; a leaf closes the run of absorbed prefixes
            members.append((attachs[src], lens[src]))
            if not plan["absorbed"][r]:
                self.inputs.append(inputs[src])
                self.attach_lists.append(members)
                members = []
        self.lcp_lens = list(plan["leaf_lcp"])
        self.lens = [len(ids) for ids in self.inputs]
        self.n_sequences = len(inputs)
        self.n_tokens = sum(lens)
        self._device_plan = plan

    def _apply_device_order(self, which: str) -> bool:
        """Apply a leaf order the device build already computed (valid only for the leaf order it was computed for)."""
        plan, self._device_plan = self._device_plan, None
        if plan is None:
            return False
        order = plan["order_" + which]
        self.inputs = [self.inputs[i] for i in order]
        self.attach_lists = [self.attach_lists[i] for i in order]
        self.lens = [self.lens[i] for i in order]
        self.lcp_lens = list(plan["lcp_" + which])
        return True

    # ---- stats -------------------------------------------------------------
    def get_stats(self, mode: str, block_size: Optional[int] = None) -> dict:
        stats = _get_stats(self.lens, self.lcp_lens, mode, block_size)
        stats["n_sequences"] = self.n_sequences
        stats["n_tokens"] = self.n_tokens
        return stats

    # ---- leaf reordering ---------------------------------------------------
    def permute(self, order: Sequence[int]) -> None:
        """Reorder leaves and recompute adjacent LCPs from the tokens (token_trie.py:90-94)."""
        self._device_plan = None
        self.inputs = [self.inputs[i] for i in order]
        self.attach_lists = [self.attach_lists[i] for i in order]
        self.lens = [self.lens[i] for i in order]
        if self.inputs and self.inputs[0].is_cuda:
            self.lcp_lens = _adjacent_lcps([], self.inputs, check_sorted=False)     # dta_lcp_adjacent, no host copy
        else:
            self.lcp_lens = _adjacent_lcps([_as_np(t) for t in self.inputs], self.inputs, check_sorted=False)

    def forward_permute(self) -> None:
        if self._apply_device_order("fwd"):
            return
        order, _, _ = CompressedTrie(self.lens, self.lcp_lens).get_order_forward()
        self.permute(order)

    def backward_permute(self) -> None:
        if self._apply_device_order("bwd"):
            return
        order, _, _ = CompressedTrie(self.lens, self.lcp_lens).get_order_backward()
        self.permute(order)

    def random_permute(self) -> None:
        order = CompressedTrie(self.lens, self.lcp_lens).get_order_random()
        self.permute(order)
