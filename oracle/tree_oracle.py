"""ORACLE (test infrastructure, never shipped, never imported by the product package).

fp32 CPU restatements of the floating-point side of the hot path:

  logprob_entropy        reference vocab_parallel.py:13-27 (log_softmax, gather, -sum p log p)
  dense_forward/backward reference dense.py:7-53 (one causal pass per sequence, no sharing)
  masked_tree_attention  softmax over ancestors, the definition the kernels implement
                         (HF models/qwen3/modeling_qwen3.py:252-291 attention with the mask of
                         SURVEY.md §9); eager, O(N^2), small inputs only
  masked_tree_attention_grads
                         the same definition with its gradients, evaluated block by block so that
                         benchmark-size tries fit (used on the GPU box, fp32, no TF32)
  tree_stack_forward/backward
                         the reference's algorithm itself — DFS over the trie with a KV stack,
                         suffix-vs-prefix causal attention per push, gradients stitched through
                         the cached K/V (reference tree_training_engine.py:146-230, 315-487,
                         515-616) — restated as a recursion over trie segments.  This is the
                         CPU baseline bench.py times.

Pinned by tests/golden (vectors produced by the unmodified reference in the build
container) and by live differential tests when /root/reference is present.
Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import trie_oracle


# ----------------------------------------------------------------------------------------
# logprob / entropy
# ----------------------------------------------------------------------------------------
def logprob_entropy(logits: torch.Tensor, labels: torch.Tensor, temperature: float = 1.0):
    lsm = F.log_softmax(logits.float() / temperature, dim=-1)
    ent = -(lsm.exp() * lsm).sum(-1)
    lp = lsm.gather(-1, labels.unsqueeze(-1)).squeeze(-1)
    return lp, ent


def example_loss(logprob: torch.Tensor, entropy: torch.Tensor, attach: dict):
    """The loss the reference's drivers use (run.py:149-152)."""
    return attach["w_logprobs"] * logprob.mean() + attach["w_entropy"] * entropy.mean()


# ----------------------------------------------------------------------------------------
# dense per-sequence baseline
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def dense_forward(model, seqs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    out = []
    for ids in seqs:
        ids = ids.unsqueeze(0).to(model.device)
        logits = model(input_ids=ids, use_cache=False).logits
        lp, _ = logprob_entropy(logits[0, :-1], ids[0, 1:])
        out.append(lp)
    return out


def dense_backward(model, seqs: Sequence[torch.Tensor], attachs: Sequence[dict], loss_fn: Callable) -> float:
    total = 0.0
    for ids, attach in zip(seqs, attachs):
        ids = ids.unsqueeze(0).to(model.device)
        logits = model(input_ids=ids, use_cache=False).logits[0]
        lp, _ = logprob_entropy(logits[:-1], ids[0, 1:])
        _, ent = logprob_entropy(logits, ids[0])          # entropy has length L, labels unused
        loss = loss_fn(lp, ent, attach)
        loss.backward()
        total += float(loss.item())
    return total


# ----------------------------------------------------------------------------------------
# masked attention definition (kernel-level oracle)
# ----------------------------------------------------------------------------------------
def masked_tree_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, subtree_end: torch.Tensor,
                          scale: float):
    """q [N,Hq,D], k/v [N,Hkv,D] (any float dtype) -> (out [N,Hq,D] fp32, lse [Hq,N] fp32).

    fp32 math on the given values; GQA by head repetition."""
    N, Hq, D = q.shape
    g = Hq // k.shape[1]
    qf = q.float().transpose(0, 1)                                   # [Hq,N,D]
    kf = k.float().repeat_interleave(g, dim=1).transpose(0, 1)
    vf = v.float().repeat_interleave(g, dim=1).transpose(0, 1)
    idx = torch.arange(N, device=q.device)
    vis = (idx[None, :] <= idx[:, None]) & (idx[:, None] < subtree_end.to(q.device)[None, :N])
    s = torch.einsum("hqd,hkd->hqk", qf, kf) * scale
    s = s.masked_fill(~vis[None], float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse[..., None])
    out = torch.einsum("hqk,hkd->hqd", p, vf).transpose(0, 1)
    return out, lse


def masked_tree_attention_grads(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dout: torch.Tensor,
                                subtree_end: torch.Tensor, scale: float, q_chunk: int = 2048):
    """(out, dq, dk, dv) in fp32 of the SAME definition as ``masked_tree_attention``, evaluated one KV head and one
    block of ``q_chunk`` queries at a time (autograd per block, gradients accumulated), so that benchmark-size tries
    (16k-67k tokens, Hq 16 / Hkv 8) fit: a block materialises group x q_chunk x N fp32 scores instead of Hq x N x N.
    Queries of block [a, b) can only see keys < b."""
    N, Hq, D = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    dev = q.device
    end = subtree_end[:N].to(dev)
    out = torch.empty((N, Hq, D), dtype=torch.float32, device=dev)
    dq = torch.empty((N, Hq, D), dtype=torch.float32, device=dev)
    dk = torch.zeros((N, Hkv, D), dtype=torch.float32, device=dev)
    dv = torch.zeros((N, Hkv, D), dtype=torch.float32, device=dev)
    for h in range(Hkv):
        for a in range(0, N, q_chunk):
            b = min(a + q_chunk, N)
            qq = q[a:b, h * g:(h + 1) * g].float().transpose(0, 1).detach().requires_grad_(True)     # [g, nq, D]
            kk = k[:b, h].float().detach().requires_grad_(True)                                       # [b, D]
            vv = v[:b, h].float().detach().requires_grad_(True)
            iq = torch.arange(a, b, device=dev)
            ik = torch.arange(b, device=dev)
            vis = (ik[None, :] <= iq[:, None]) & (iq[:, None] < end[None, :b])
            sc = torch.einsum("gqd,kd->gqk", qq, kk) * scale
            sc = sc.masked_fill(~vis[None], float("-inf"))
            lse = torch.logsumexp(sc, dim=-1)
            o = torch.einsum("gqk,kd->gqd", torch.exp(sc - lse[..., None]), vv)
            o.backward(dout[a:b, h * g:(h + 1) * g].float().transpose(0, 1))
            out[a:b, h * g:(h + 1) * g] = o.detach().transpose(0, 1)
            dq[a:b, h * g:(h + 1) * g] = qq.grad.transpose(0, 1)
            dk[:b, h] += kk.grad
            dv[:b, h] += vv.grad
            del sc, o, lse, vis
    return out, dq, dk, dv


# ----------------------------------------------------------------------------------------
# the reference algorithm: DFS over trie segments with a KV stack
# ----------------------------------------------------------------------------------------
class _Seg:
    """One compressed-trie edge: tokens[start:end) of leaf `leaf`, children below it."""
    __slots__ = ("leaf", "start", "end", "kids", "ends_here")

    def __init__(self, leaf, start, end):
        self.leaf, self.start, self.end = leaf, start, end
        self.kids: List["_Seg"] = []
        self.ends_here: List[Tuple[dict, int]] = []     # sequences whose last token lies in this segment


def _segments(leaves: List[torch.Tensor], lcps: Sequence[int], attach_lists) -> List[_Seg]:
    """Cut the DFS-ordered leaves into compressed-trie segments (forest roots returned)."""
    roots: List[_Seg] = []
    spine: List[_Seg] = []
    for i, ids in enumerate(leaves):
        f = lcps[i - 1] if i else 0
        while spine and spine[-1].start >= f:
            spine.pop()
        if spine and spine[-1].end > f:
            top = spine[-1]                     # the fork falls inside this segment: split it at f
            low = _Seg(top.leaf, f, top.end)
            low.kids, top.kids = top.kids, [low]
            low.ends_here = [(a, L) for a, L in top.ends_here if L > f]
            top.ends_here = [(a, L) for a, L in top.ends_here if L <= f]
            top.end = f
        seg = _Seg(i, f, int(ids.numel()))
        for a, L in attach_lists[i]:
            if L > f:
                seg.ends_here.append((a, L))
            else:                               # a merged prefix that ends above the fork
                for anc in spine:
                    if anc.start < L <= anc.end:
                        anc.ends_here.append((a, L))
                        break
        (spine[-1].kids if spine else roots).append(seg)
        spine.append(seg)
    return roots


def _run_segment(model, tokens: torch.Tensor, past_kv, position0: int, grad: bool):
    """HF causal forward of a suffix against the prefix cache; returns (logits, new K, new V lists)."""
    from transformers.cache_utils import DynamicCache
    cache = DynamicCache()
    if past_kv is not None:
        for l, (k, v) in enumerate(past_kv):
            cache.update(k, v, layer_idx=l)
    ctx = torch.enable_grad() if grad else torch.no_grad()
    with ctx:
        out = model(tokens.unsqueeze(0), past_key_values=cache, use_cache=True)
    ks = [layer.keys[:, :, position0:, :] for layer in out.past_key_values.layers]
    vs = [layer.values[:, :, position0:, :] for layer in out.past_key_values.layers]
    return out.logits[0], ks, vs


def tree_stack_forward(model, leaves, lcps, attach_lists, n_sequences: int) -> List[torch.Tensor]:
    """Forward-only DFS with a KV stack (reference tree_training_engine.py:146-230, 515-553)."""
    n_layers = model.config.num_hidden_layers
    returns: List[Optional[torch.Tensor]] = [None] * n_sequences
    kv_stack: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None
    lp_stack = torch.zeros(0)
    last_logits_at: Dict[int, torch.Tensor] = {}
    for i, ids in enumerate(leaves):
        s = lcps[i - 1] if i else 0
        if kv_stack is not None:
            kv_stack = [(k[:, :, :s], v[:, :, :s]) for k, v in kv_stack]
        new = ids[s:]
        logits, ks, vs = _run_segment(model, new, kv_stack if s > 0 else None, s, grad=False)
        lp_new, _ = logprob_entropy(logits[:-1], new[1:])
        lp_stack = lp_stack[:max(s - 1, 0)]
        if s > 0:
            edge = F.log_softmax(last_logits_at[s - 1].float(), -1)[new[0]].reshape(1)
            lp_stack = torch.cat([lp_stack, edge, lp_new])
        else:
            lp_stack = lp_new
        if kv_stack is None or s == 0:
            kv_stack = list(zip(ks, vs))
        else:
            kv_stack = [(torch.cat([k0, k1], 2), torch.cat([v0, v1], 2)) for (k0, v0), k1, v1 in zip(kv_stack, ks, vs)]
        for d in range(s, int(ids.numel())):
            last_logits_at[d] = logits[d - s]
        for attach, L in attach_lists[i]:
            returns[attach["_sequence_batch_id"]] = lp_stack[:L - 1].clone()
    return returns


def tree_stack_backward(model, leaves, lcps, attach_lists, loss_fn: Callable) -> float:
    """Backward DFS with prefix sharing (reference tree_training_engine.py:315-487, 555-616).

    Each trie segment is run ONCE with a graph, against the (graph-carrying) K/V of its
    ancestors; children are processed depth-first and their losses added; one backward per
    root frees the graph.  Same arithmetic as the reference's push/pop with injected KV
    gradients: shared-prefix tokens are computed once and receive the summed gradient of
    every branch below them.
    """
    roots = _segments(list(leaves), list(lcps), attach_lists)
    total = 0.0

    def visit(seg: _Seg, past_kv, lp_prefix, ent_prefix, edge_logits):
        ids = leaves[seg.leaf][seg.start:seg.end]
        logits, ks, vs = _run_segment(model, ids, past_kv, seg.start, grad=True)
        lp_in, ent = logprob_entropy(logits, torch.cat([ids[1:], ids[:1]]))       # last label is a dummy
        lp_in = lp_in[:-1]
        if edge_logits is not None:
            edge = F.log_softmax(edge_logits.float(), -1)[ids[0]].reshape(1)
            lp_all = torch.cat([lp_prefix, edge, lp_in])
        else:
            lp_all = lp_in
        ent_all = torch.cat([ent_prefix, ent]) if ent_prefix is not None else ent
        loss = None
        for attach, L in seg.ends_here:
            term = loss_fn(lp_all[:L - 1], ent_all[:L], attach)
            loss = term if loss is None else loss + term
        if past_kv is None:
            kv = list(zip(ks, vs))
        else:
            kv = [(torch.cat([k0, k1], 2), torch.cat([v0, v1], 2)) for (k0, v0), k1, v1 in zip(past_kv, ks, vs)]
        for child in seg.kids:
            sub = visit(child, kv, lp_all, ent_all, logits[-1])
            if sub is not None:
                loss = sub if loss is None else loss + sub
        return loss

    for root in roots:
        loss = visit(root, None, torch.zeros(0), None, None)
        if loss is not None:
            loss.backward()
            total += float(loss.item())
    return total
