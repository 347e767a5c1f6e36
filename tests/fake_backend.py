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


# ---- torch restatements of the six kernel wrappers of dynamictreeattn_b200/vocab_parallel.py (contracts:
# include/dta_b200.h, "Log-prob + entropy over a vocab shard") so that the vocab-parallel HOST logic -- block
# packing, the single all-gather, the merge, the collective-free backward -- runs under gloo on the CPU ------------
def k_row_stats(x, inv_temp, labels, vocab_start, stats):
    z = x.float() * inv_temp
    m = z.max(dim=-1).values
    e = torch.exp(z - m[:, None])
    stats[:, 0], stats[:, 1], stats[:, 2] = m, e.sum(-1), (z * e).sum(-1)
    stats[:, 3] = 0.0
    if labels is not None:
        c = labels.long() - vocab_start
        ok = (c >= 0) & (c < x.shape[1])
        rows = torch.nonzero(ok).squeeze(1)
        stats[rows, 3] = z[rows, c[rows]]


def k_gather_labels(x, inv_temp, labels, row_of, vocab_start, out):
    c = labels.long() - vocab_start
    r = row_of.long()
    ok = (c >= 0) & (c < x.shape[1]) & (r >= 0)
    out.zero_()
    idx = torch.nonzero(ok).squeeze(1)
    out[idx] = x.float()[r[idx], c[idx]] * inv_temp


def k_finalize(blocks, n_rows, want_lp):
    st = blocks[:, :4 * n_rows].reshape(blocks.shape[0], n_rows, 4)
    m = st[:, :, 0].max(dim=0).values
    f = torch.exp(st[:, :, 0] - m[None])
    s = (st[:, :, 1] * f).sum(0)
    t = (st[:, :, 2] * f).sum(0)
    logz = m + torch.log(s)
    ent = logz - t / s
    lp = st[:, :, 3].sum(0) - logz if want_lp else None
    return logz, ent, lp


def k_backward(x, inv_temp, logz, entropy, g_ent, g_rowsum, row_labels, vocab_start, out):
    z = x.float() * inv_temp
    lp = z - logz[:, None]
    p = lp.exp()
    R = x.shape[0]
    a = g_ent.float() if g_ent is not None else torch.zeros(R)
    G = g_rowsum.float() if g_rowsum is not None else torch.zeros(R)
    d = inv_temp * p * (a[:, None] * (-entropy[:, None] - lp) - G[:, None])
    if row_labels is not None:
        c = row_labels.long() - vocab_start
        rows = torch.nonzero((c >= 0) & (c < x.shape[1])).squeeze(1)
        d[rows, c[rows]] += inv_temp * G[rows]
    out.copy_(d.to(out.dtype))


def k_rowsum(g_lp, row_of, n_rows):
    out = torch.zeros(n_rows, dtype=torch.float32)
    ok = row_of >= 0
    out.index_add_(0, row_of[ok].long(), g_lp[ok].float())
    return out


def k_scatter(out, inv_temp, labels, row_of, g_lp, vocab_start):
    c = labels.long() - vocab_start
    r = row_of.long()
    idx = torch.nonzero((c >= 0) & (c < out.shape[1]) & (r >= 0)).squeeze(1)
    out[r[idx], c[idx]] += (inv_temp * g_lp[idx]).to(out.dtype)


def install_vocab_parallel(monkeypatch):
    """Swap the C-ABI wrappers of vocab_parallel.py for the restatements above (CPU tensors allowed)."""
    import dynamictreeattn_b200._cabi as cabi
    import dynamictreeattn_b200.vocab_parallel as vp
    for name, fn in (("_k_row_stats", k_row_stats), ("_k_gather_labels", k_gather_labels), ("_k_finalize", k_finalize),
                     ("_k_backward", k_backward), ("_k_rowsum", k_rowsum), ("_k_scatter", k_scatter)):
        monkeypatch.setattr(vp, name, fn)
    monkeypatch.setattr(cabi, "require_cuda", lambda *a: None)


def install(monkeypatch):
    import dynamictreeattn_b200.attention as attention
    import dynamictreeattn_b200.tree_training_engine as engine
    monkeypatch.setattr(engine, "linearize_trie", linearize_trie)
    monkeypatch.setattr(attention, "tree_attention", tree_attention)
    monkeypatch.setattr(engine, "logprob_entropy_rows", logprob_entropy_rows)
    monkeypatch.setattr(engine, "logprob_entropy_rows_backward", logprob_entropy_rows_backward)
