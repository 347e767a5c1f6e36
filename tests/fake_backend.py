"""TEST-ONLY stand-ins for the CUDA ops, built from the oracle.

The product has no CPU path.  To exercise the engine's HOST logic (edge plan, per-sequence
path gathers, head chunking, autograd wiring, HF plugin) in the GPU-less CI run, these
functions are monkey-patched over the C-ABI wrappers by the ``cpu_backend`` fixture.
They are plain torch/numpy restatements of the kernel contracts in include/dta_b200.h.
"""
from __future__ import annotations

import numpy as np
import torch

from dynamictreeattn_b200.linearize import TILE, TrieLayout
from oracle import trie_oracle, tree_oracle


def linearize_trie(token_trie, device) -> TrieLayout:
    lin = trie_oracle.linearize([t.tolist() for t in token_trie.inputs], token_trie.lcp_lens)
    N = len(lin["tok"])
    n_pad = (N + TILE - 1) // TILE * TILE
    dev = torch.device(device)

    def pad(a, fill):
        out = np.full(n_pad, fill, dtype=np.int32)
        out[:N] = a
        return torch.from_numpy(out).to(dev)

    meta = trie_oracle.tile_meta(lin["subtree_end"], TILE)
    path_off = [0]
    for p in lin["paths"]:
        path_off.append(path_off[-1] + len(p))
    t32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int32)).to(dev)
    pairs = int((lin["depth"] + 1).sum())
    return TrieLayout(
        n_nodes=N, n_pad=n_pad, n_leaf=len(token_trie.lens), lens=list(token_trie.lens),
        lcp_lens=list(token_trie.lcp_lens), tok=torch.from_numpy(lin["tok"]).to(dev),
        depth=pad(lin["depth"], 0), parent=pad(lin["parent"], -1), subtree_end=pad(lin["subtree_end"], 0),
        leaf_of=pad(lin["leaf_of"], -1), leaf_off=t32(lin["leaf_off"]),
        path=t32(np.concatenate(lin["paths"])), path_off=path_off,
        kmax_end=t32(meta["kmax_end"]), kmin_end=t32(meta["kmin_end"]), q_order=t32(meta["q_order"]),
        q_nlive=t32(meta["q_nlive"]), k_order=t32(meta["k_order"]), k_qlast=t32(meta["k_qlast"]),
        pair_count=torch.tensor([pairs], dtype=torch.int64, device=dev))


def tree_attention(q, k, v, layout: TrieLayout, scale=None):
    if scale is None:
        scale = q.shape[-1] ** -0.5
    out, _ = tree_oracle.masked_tree_attention(q, k, v, layout.subtree_end[:layout.n_nodes].long(), scale)
    return out.to(q.dtype)


def logprob_entropy_rows(x, labels, row_of, inv_temp=1.0, tp_group=None):
    z = x.float() * inv_temp
    logz = torch.logsumexp(z, dim=-1)
    p = torch.exp(z - logz[:, None])
    ent = logz - (p * z).sum(-1)
    x_label = z[row_of.long(), labels.long()]
    return x_label, logz, ent


def logprob_entropy_rows_backward(x, labels, row_of, inv_temp, vocab_start, logz, entropy, g_lp, g_ent, out=None):
    z = x.float() * inv_temp
    lp = z - logz[:, None]
    p = lp.exp()
    R = x.shape[0]
    G = torch.zeros(R, dtype=torch.float32, device=x.device)
    if g_lp is not None:
        G.index_add_(0, row_of.long(), g_lp.float())
    a = g_ent.float() if g_ent is not None else torch.zeros(R, device=x.device)
    d = inv_temp * p * (a[:, None] * (-entropy[:, None] - lp) - G[:, None])
    if g_lp is not None:
        d[row_of.long(), labels.long() - vocab_start] += inv_temp * g_lp.float()
    d = d.to(x.dtype)
    if out is not None:
        out.copy_(d)
        return out
    return d


def install(monkeypatch):
    import dynamictreeattn_b200.attention as attention
    import dynamictreeattn_b200.tree_training_engine as engine
    monkeypatch.setattr(engine, "linearize_trie", linearize_trie)
    monkeypatch.setattr(attention, "tree_attention", tree_attention)
    monkeypatch.setattr(engine, "logprob_entropy_rows", logprob_entropy_rows)
    monkeypatch.setattr(engine, "logprob_entropy_rows_backward", logprob_entropy_rows_backward)
