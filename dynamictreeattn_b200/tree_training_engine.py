"""TreeTrainingEngine — forward / backward over a TokenTrie in ONE linearised pass.

Drop-in for the reference class (tree_training_engine.py:52-616): same constructor,
``forward(model, token_trie) -> List[Tensor]`` and
``backward(model, token_trie, loss_fn, block_size, cut_f1_tail) -> float`` with the
same conventions (per-sequence fp32 log-probs of length ``len(seq)-1``; gradients are
ACCUMULATED into ``param.grad``; the returned loss is a Python float).

What runs is different.  The reference walks the trie with a push/pop KV stack, calls
the HF model once or twice per leaf and stitches the backward from injected gradients
(tree_training_engine.py:146-313, 315-487, 555-616).  Here the trie is linearised once
on the GPU (``linearize.py``: parent / depth / ancestor intervals), the model trunk runs
ONCE over all N trie tokens with ``position_ids = depth`` and the ancestor-masked tcgen05
attention kernels, and the vocabulary head is evaluated in row chunks by the fused
log-prob/entropy kernels with per-EDGE labels (a fork node has one label per child,
which is why the reference keeps whole logits rows at fork positions,
tree_training_engine.py:217-220, 369-372).  The result is mathematically the reference's
 and independent of ``block_size`` / ``cut_f1_tail`` / leaf order,
which are accepted for compatibility: ``block_size`` only sizes the head's row chunks.

Memory.  The reference's footprint is bounded by the stack depth; one linearised pass keeps the activations of
all N trie tokens (~1.4 MB per token for Qwen3-0.6B).  A trie over the engine's token budget
(``max_trie_tokens``; by default derived from free HBM) is therefore cut into contiguous DFS ranges of sequences
with the reference's own partitioner (``LB_by_DFS_and_TM``) and the ranges run one after the other: losses add,
gradients accumulate, the shared prefix of a cut is recomputed once per range — same result, bounded memory.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import torch

from . import _cabi
from .attention import use_tree_attention
from .fused_layers import fused_trunk_forward, supports_fused_trunk
from .linearize import TrieLayout, linearize_trie
from .vocab_parallel import (fused_head_supported, lm_head_dlogits_rows, lm_head_logits_and_stats_rows,
                             lm_head_logprob_entropy_rows, logprob_entropy_rows, logprob_entropy_rows_backward)

__all__ = ["TreeTrainingEngine", "_get_forkpos"]


def _get_forkpos(lens, lcp_lens, block_size) -> list:
    """Stack positions whose full logits row the REFERENCE engine must retain
    (tree_training_engine.py:12-49): ``lcp-1`` of every fork plus ``pop_start-1`` of every
    block boundary.  The single-pass engine needs no such bookkeeping (every row's labels
    are known up front); kept for API parity and for the cost-model tests."""
    out = set()
    for lcp in lcp_lens:
        if lcp > 0:
            out.add(lcp - 1)
    if block_size is not None:
        for i, end in enumerate(lens):
            start = 0 if i == len(lcp_lens) else lcp_lens[i]
            n = end - start
            n_blocks = -(-n // block_size)
            blk = -(-n // n_blocks)
            for b in range(n_blocks):
                cut = max(end - (b + 1) * blk, start)
                if cut > 0:
                    out.add(cut - 1)
    return sorted(out)


def _decoder_and_head(model):
    """(trunk returning last_hidden_state, lm_head weight [V, H])."""
    trunk = getattr(model, "model", None)
    if trunk is None and hasattr(model, "get_decoder"):
        trunk = model.get_decoder()
    head = model.get_output_embeddings() if hasattr(model, "get_output_embeddings") else model.lm_head
    if trunk is None or head is None:
        raise _cabi.DtaError("model must expose .model (decoder trunk) and an lm_head")
    if getattr(head, "bias", None) is not None:
        raise _cabi.DtaError("lm_head with bias is not supported")
    return trunk, head.weight


_MM_OUT_DTYPE = None


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with fp32 output (one rounding at the very end of the dW accumulation); falls back
    to a bf16-output GEMM per chunk on torch builds without ``out_dtype``."""
    global _MM_OUT_DTYPE
    if _MM_OUT_DTYPE is None:
        try:
            torch.mm(a[:1], b[:, :1], out_dtype=torch.float32)
            _MM_OUT_DTYPE = True
        except (TypeError, RuntimeError):
            _MM_OUT_DTYPE = False
    if _MM_OUT_DTYPE:
        return torch.mm(a, b, out_dtype=torch.float32)
    return torch.mm(a, b).float()


_ADDMM_OUT_DTYPE = None


def _addmm_f32_(acc: torch.Tensor, a: torch.Tensor, b: torch.Tensor) -> None:
    """acc (fp32) += a @ b (bf16 operands) inside the GEMM's epilogue when the torch build can (``addmm`` with
    ``out_dtype``): saves reading and writing the 0.6 GB fp32 dW accumulator once more per row chunk; otherwise an
    fp32-output GEMM followed by an add."""
    global _ADDMM_OUT_DTYPE
    if _ADDMM_OUT_DTYPE is None:
        try:
            t = torch.zeros((1, 1), dtype=torch.float32, device=a.device)
            torch.addmm(t, a[:1, :8], b[:8, :1], out_dtype=torch.float32, out=t)
            _ADDMM_OUT_DTYPE = True
        except (TypeError, RuntimeError):
            _ADDMM_OUT_DTYPE = False
    if _ADDMM_OUT_DTYPE:
        torch.addmm(acc, a, b, out_dtype=torch.float32, out=acc)
    else:
        acc.add_(_mm_f32(a, b))


class _EdgePlan:
    """Edges (parent row -> child node) grouped by parent row, so a chunk of logits rows
    maps to one contiguous run of labels."""

    def __init__(self, layout: TrieLayout, chunk_rows: int):
        N = layout.n_nodes
        parent = layout.parent[:N]
        child = torch.nonzero(parent >= 0).squeeze(1)
        rows = parent[child].to(torch.int64)
        order = torch.argsort(rows, stable=True)
        self.child = child[order]                      # node that receives the log-prob
        self.row = rows[order]                         # logits row the label is read from
        self.label = layout.tok[self.child]            # token of the child
        self.n_nodes = N
        self.chunk_rows = chunk_rows
        # edge range of every row chunk, ONE device read per step (not one per chunk and pass)
        cuts = torch.arange(0, N + chunk_rows, chunk_rows, device=self.row.device).clamp_(max=N)
        self.bounds = torch.searchsorted(self.row, cuts).tolist()

    def span(self, a: int, b: int) -> Tuple[int, int]:
        """Edges whose label row lies in [a, b); (a, b) must be a chunk of the plan."""
        i = a // self.chunk_rows
        assert a == i * self.chunk_rows and b == min(a + self.chunk_rows, self.n_nodes)
        return self.bounds[i], self.bounds[i + 1]


class _TreeHeadLoss(torch.autograd.Function):
    """loss(hidden, W) = sum_seq loss_fn(lp[path[1:L]], H[path[:L]], attach).

    Forward evaluates the head in row chunks (logits = hidden W^T, fused stats kernel),
    runs the user's ``loss_fn`` on the small per-node vectors, and — because the head is
    the last op before the scalar loss — already produces d loss / d hidden and d loss / dW
    chunk by chunk (logits recomputed, gradient written in place over them), so no
    ``[N, V]`` tensor outlives a chunk.  Backward just scales the stashed gradients.
    """

    @staticmethod
    def forward(ctx, hidden: torch.Tensor, weight: torch.Tensor, engine, layout: TrieLayout, token_trie, loss_fn,
                chunk_rows: int):
        N = layout.n_nodes
        h = hidden.detach()
        W = weight.detach()
        need_h = hidden.requires_grad
        need_w = weight.requires_grad
        fused = engine._use_fused_head(h, W, training=need_h or need_w)
        if fused:
            # N1: lm_head fused into the statistics (csrc/head_fwd.cu) -- no logits in HBM, in either pass
            lp_nodes, ent_nodes, logz, plan = engine._head_forward_fused(h, W, layout, chunk_rows)
            kept = None
        else:
            # With HBM to spare (180 GB on B200) the bf16 logits chunks stay resident between the statistics
            # pass and the gradient pass, which saves re-running the [rows, hidden] x [hidden, vocab] GEMM.
            keep = (need_h or need_w) and engine._can_keep_logits(N, W.shape[0], h.element_size(), h.device)
            lp_nodes, ent_nodes, logz, plan, kept = engine._head_forward(h, W, layout, chunk_rows, keep_logits=keep)

        # ---- user loss on per-node vectors (tiny autograd graph) --------------------------
        with torch.enable_grad():
            lp_leaf = lp_nodes.requires_grad_(True)
            ent_leaf = ent_nodes.clone().requires_grad_(True)
            total = None
            for i, attach_list in enumerate(token_trie.attach_lists):
                for attach, length in attach_list:
                    path = layout.path_of_leaf(i, length).long()
                    term = loss_fn(lp_leaf[path[1:]], ent_leaf[path], attach)
                    total = term if total is None else total + term
            if total is None:
                total = lp_leaf.sum() * 0.0
            g_lp, g_ent = torch.autograd.grad(total, [lp_leaf, ent_leaf], allow_unused=True)
        loss = total.detach()

        # ---- head backward, chunked -------------------------------------------------------
        dh = torch.empty_like(h) if need_h else None
        dW32 = torch.zeros(W.shape, dtype=torch.float32, device=W.device) if need_w else None
        if need_h or need_w:
            g_lp = torch.zeros_like(lp_nodes) if g_lp is None else g_lp.float()
            g_ent = torch.zeros_like(ent_nodes) if g_ent is None else g_ent.float()
            dl_buf = torch.empty((min(chunk_rows, N), W.shape[0]), dtype=h.dtype, device=h.device) if fused else None
            for ci, a in enumerate(range(0, N, chunk_rows)):
                b = min(a + chunk_rows, N)
                e0, e1 = plan.span(a, b)
                row_of = (plan.row[e0:e1] - a).to(torch.int32)
                if fused:           # recompute GEMM + log-prob backward in one kernel, gradient chunk in bf16
                    logits = lm_head_dlogits_rows(h[a:b], W, plan.label[e0:e1], row_of, 1.0, logz[a:b], ent_nodes[a:b],
                                                  g_lp[plan.child[e0:e1]], g_ent[a:b], out=dl_buf[:b - a])
                else:
                    if kept is not None:
                        logits, kept[ci] = kept[ci], None
                    else:
                        logits = h[a:b] @ W.t()
                    logprob_entropy_rows_backward(logits, plan.label[e0:e1], row_of, 1.0, 0, logz[a:b], ent_nodes[a:b],
                                                  g_lp[plan.child[e0:e1]], g_ent[a:b], out=logits)
                if need_h:
                    torch.mm(logits, W, out=dh[a:b])
                if need_w:
                    _addmm_f32_(dW32, logits.t(), h[a:b])
                del logits
        ctx.dh = dh
        ctx.dW = dW32.to(W.dtype) if need_w else None
        return loss

    @staticmethod
    def backward(ctx, g):
        dh = ctx.dh * g.to(ctx.dh.dtype) if ctx.dh is not None else None
        dW = ctx.dW * g.to(ctx.dW.dtype) if ctx.dW is not None else None
        ctx.dh = ctx.dW = None
        return dh, dW, None, None, None, None, None


class TreeTrainingEngine:
    """Engine for forward / backward computation over sequences with shared prefixes."""

    def __init__(self, model_config, device, dtype: torch.dtype, max_seq_len: int, forward_only: bool = False):
        self.model = None
        self.model_config = model_config
        self.device = torch.device(device)
        self.dtype = dtype
        self.max_seq_len = max_seq_len
        self.forward_only = forward_only
        # fields the reference reads (tree_training_engine.py:102-106); validated early
        self.n_layers = model_config.num_hidden_layers
        self.n_kv_heads = model_config.num_key_value_heads
        self.head_dim = model_config.head_dim if hasattr(model_config, "head_dim") and model_config.head_dim \
            else model_config.hidden_size // model_config.num_attention_heads
        self.head_chunk_rows = 8192
        self.keep_head_logits = True     # (unfused head only) keep bf16 logits between the head's two passes when HBM allows
        # N1: lm_head fused into the log-prob statistics and their backward (csrc/head_fwd.cu: no logits in HBM).
        # True / False force it.  None = by measured cost (profiles/r02/head_bench.log, visit12 in-step A/B): the fused
        # FORWARD beats GEMM + statistics (14.8 vs 21.6 ms at 67k x 151,936) and runs whenever no backward follows
        # (engine.forward); a training step's backward has to recompute the GEMM without kept logits (2.3 ms per
        # 8192-row chunk fused, against 1.3 ms on kept logits), which makes the no-logits step 1 % slower and 17 GB
        # smaller -- so training goes logit-free when the logits would not fit anyway (_can_keep_logits); while they
        # fit, the SAME fused launch also writes them (dta_lm_head_stats_logits) and the backward runs on kept logits.
        env = os.environ.get("DTA_FUSED_HEAD")
        self.fused_head: Optional[bool] = None if env is None else env not in ("0", "")
        self.use_fused_trunk = True      # Qwen3-style bf16 models; others go through the HF attention plugin
        self.max_trie_tokens: Optional[int] = None    # None: derive from free HBM (CUDA) / unlimited (elsewhere)
        # Activation recomputation (fused trunk only).  None = automatic: a trie whose full activations exceed the
        # token budget first switches to per-layer recomputation (fused_layers._RecomputedLayer: 4-6x less memory,
        # ~7 % more time) and only a trie that does not fit even then is cut into DFS ranges.  True / False force it.
        self.recompute: Optional[bool] = None
        self.last_recompute = False
        self.last_layout: Optional[TrieLayout] = None
        self.last_n_ranges = 1           # sub-tries the last forward/backward call ran as
        self.pass_count_listener = None  # callable(n_passes): told how many backward passes a step runs as (DP reducer)
        self._budget_cache = {}

    # ---- memory-bounded execution ------------------------------------------------------------
    def _token_budget(self, training: bool, recompute: bool = False) -> Optional[int]:
        """Trie tokens one linearised pass may hold.  Saved activations per token and layer (bf16): 7 hidden-wide
        rows, q/k/v and their normed copies, the attention output, three MLP-wide rows; the backward adds fp32
        dq/dk/dv accumulators and gradients of the same shapes.  With ``recompute`` a layer keeps its two input rows
        and the attention output; three layers' worth of full activations are live at a time (the layer being
        recomputed, its gradients, allocator slack)."""
        if self.max_trie_tokens is not None:
            return int(self.max_trie_tokens)              # an explicit cap is a cap on tokens per pass, whatever the mode
        if self.device.type != "cuda":
            return None
        key = (id(self.model), training, recompute)
        if key in self._budget_cache:                    # cudaMemGetInfo once per model, not once per step
            return self._budget_cache[key]
        cfg = self.model_config
        hq = cfg.num_attention_heads
        inter = getattr(cfg, "intermediate_size", 4 * cfg.hidden_size)
        per_layer = 7 * cfg.hidden_size + (3 * hq + 3 * self.n_kv_heads) * self.head_dim + 3 * inter
        if not training:
            per_token = 2 * per_layer * 2
        elif recompute:
            per_token = 2 * ((2 * cfg.hidden_size + hq * self.head_dim) * self.n_layers + 3 * per_layer)
        else:
            per_token = 2 * per_layer * self.n_layers
        per_token += 16 * (hq + self.n_kv_heads) * self.head_dim
        free, _ = torch.cuda.mem_get_info(self.device)
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        usable = free - (16 << 30)                       # head chunks, workspaces, allocator slack
        self._budget_cache[key] = max(int(0.85 * usable / per_token), 4096)
        return self._budget_cache[key]

    @staticmethod
    def _sequences_of(token_trie):
        """(tokens, attach) of every original sequence of a trie (leaf i holds its folded prefixes)."""
        out = []
        for ids, attach_list in zip(token_trie.inputs, token_trie.attach_lists):
            for attach, length in attach_list:
                out.append((ids[:length], attach))
        return out

    def _ranges(self, token_trie, budget: int):
        """Cut a trie that exceeds ``budget`` trie tokens into sub-tries of contiguous DFS ranges; None if it fits
        (or cannot be cut).  Attach dicts are shallow copies that keep the ORIGINAL ``_sequence_batch_id``."""
        from .data_parallel import partition
        from .token_trie import TokenTrie
        n_tok = sum(token_trie.lens) - sum(token_trie.lcp_lens)      # trie tokens (trie.py:8-13)
        if n_tok <= budget or token_trie.n_sequences < 2:
            return None
        seqs = self._sequences_of(token_trie)
        toks = [t for t, _ in seqs]
        k = min(len(seqs), -(-n_tok // budget))
        while True:
            bins = [b for b in partition(toks, k, "LB_by_DFS_and_TM", None, "backward", None) if b]
            tries = []
            for b in bins:
                att = [dict(seqs[i][1]) for i in b]
                t = TokenTrie([toks[i] for i in b], att)
                for a, i in zip(att, b):
                    a["_sequence_batch_id"] = seqs[i][1].get("_sequence_batch_id", i)
                tries.append(t)
            worst = max(t.get_stats("forward")["n_tree_tokens"] for t in tries)
            if worst <= budget or k >= len(seqs):
                return tries
            k = min(len(seqs), max(k + 1, -(-k * worst // budget)))

    # ---- shared pieces ---------------------------------------------------------------------
    def _check_lengths(self, token_trie) -> None:
        for ids in token_trie.inputs:
            n = int(ids.numel())
            assert n <= self.max_seq_len, (
                f"Exceeds max_seq_len: cur_len=0, new_tokens={n}, max={self.max_seq_len}")

    def _plan_memory(self, model, n_tok: int, training: bool) -> Tuple[Optional[int], bool]:
        """(token budget of one pass, recompute layers?) for a trie of ``n_tok`` tokens."""
        budget = self._token_budget(training)
        can = training and self.use_fused_trunk and supports_fused_trunk(model)
        if self.recompute is not None:
            rec = bool(self.recompute) and can
        else:
            rec = can and self.max_trie_tokens is None and budget is not None and n_tok > budget
        if rec:
            budget = self._token_budget(training, recompute=True)
        return budget, rec

    def _trunk(self, model, layout: TrieLayout, recompute: bool = False) -> torch.Tensor:
        if self.use_fused_trunk and supports_fused_trunk(model):
            return fused_trunk_forward(model, layout, recompute)   # fused sm_100a layer kernels (N3)
        trunk, _ = _decoder_and_head(model)
        N = layout.n_nodes
        with use_tree_attention(model, layout):
            out = trunk(input_ids=layout.tok.unsqueeze(0),
                        position_ids=layout.depth[:N].to(torch.int64).unsqueeze(0),
                        use_cache=False)
        return out.last_hidden_state[0]                       # [N, hidden]

    def _can_keep_logits(self, n_rows: int, vocab: int, elem: int, dev) -> bool:
        """True when all logits chunks of one step fit in free HBM with a 24 GB margin."""
        if not self.keep_head_logits or torch.device(dev).type != "cuda":
            return False
        free, _ = torch.cuda.mem_get_info(dev)
        free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        return n_rows * vocab * elem + (24 << 30) < free

    def _use_fused_head(self, h: torch.Tensor, W: torch.Tensor, training: bool = False) -> bool:
        if self.fused_head is False:
            return False
        ok = fused_head_supported(h, W)
        if self.fused_head and not ok:
            raise _cabi.DtaError("fused_head=True but hidden / lm_head weight are not bf16 CUDA tensors the kernel can address")
        if self.fused_head is None and training and ok:
            return not self._can_keep_logits(h.shape[0], W.shape[0], h.element_size(), h.device)
        return ok

    def _head_forward_fused(self, h: torch.Tensor, W: torch.Tensor, layout: TrieLayout, chunk_rows: int):
        """``_head_forward`` through the fused lm_head kernel: ONE launch over all rows, no logits."""
        N = layout.n_nodes
        plan = _EdgePlan(layout, chunk_rows)
        x_label, logz, ent_nodes = lm_head_logprob_entropy_rows(h, W, plan.label, plan.row.to(torch.int32), 1.0)
        lp_nodes = torch.zeros(N, dtype=torch.float32, device=h.device)
        lp_nodes[plan.child] = x_label - logz[plan.row]
        return lp_nodes, ent_nodes, logz, plan

    def _head_forward(self, h: torch.Tensor, W: torch.Tensor, layout: TrieLayout, chunk_rows: int,
                      keep_logits: bool = False):
        """Per-node log-prob of the edge INTO the node (roots: 0) and per-node entropy
        (+ the logits chunks themselves when ``keep_logits``)."""
        N = layout.n_nodes
        dev = h.device
        plan = _EdgePlan(layout, chunk_rows)
        lp_nodes = torch.zeros(N, dtype=torch.float32, device=dev)
        ent_nodes = torch.empty(N, dtype=torch.float32, device=dev)
        logz = torch.empty(N, dtype=torch.float32, device=dev)
        kept = [] if keep_logits else None
        # kept logits come out of the fused lm_head launch (statistics in the GEMM's epilogue, logits written once and
        # not read back in the forward); a head that is not fusable runs the library GEMM + the row kernel
        fused_keep = keep_logits and self.fused_head is not False and fused_head_supported(h, W)
        for a in range(0, N, chunk_rows):
            b = min(a + chunk_rows, N)
            e0, e1 = plan.span(a, b)
            row_of = (plan.row[e0:e1] - a).to(torch.int32)
            if fused_keep:
                logits, x_label, lz, ent = lm_head_logits_and_stats_rows(h[a:b], W, plan.label[e0:e1], row_of, 1.0)
            else:
                logits = h[a:b] @ W.t()
                x_label, lz, ent = logprob_entropy_rows(logits, plan.label[e0:e1], row_of, 1.0)
            lp_nodes[plan.child[e0:e1]] = x_label - lz[row_of.long()]
            ent_nodes[a:b] = ent
            logz[a:b] = lz
            if kept is not None:
                kept.append(logits)
            del logits
        return lp_nodes, ent_nodes, logz, plan, kept

    def _chunk_rows(self, block_size: Optional[int]) -> int:
        rows = self.head_chunk_rows
        if block_size is not None and block_size > 0:
            rows = max(rows, int(block_size))
        return rows

    # ---- public API ------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, model, token_trie) -> List[torch.Tensor]:
        """Per-sequence log-probs, indexed by ``attach['_sequence_batch_id']``
        (reference tree_training_engine.py:515-553)."""
        self.model = model
        self._check_lengths(token_trie)
        returns: List[Optional[torch.Tensor]] = [None] * token_trie.n_sequences
        if not token_trie.inputs:
            return returns
        budget = self._token_budget(training=False)
        ranges = self._ranges(token_trie, budget) if budget is not None else None
        self.last_n_ranges = len(ranges) if ranges else 1
        if ranges:
            for sub in ranges:
                sub.forward_permute()
                for sid, lp in self._forward_one(model, sub).items():
                    returns[sid] = lp
            return returns
        for sid, lp in self._forward_one(model, token_trie).items():
            returns[sid] = lp
        return returns

    def _forward_one(self, model, token_trie) -> dict:
        """{_sequence_batch_id: log-probs} of one trie that fits the budget."""
        returns = {}
        layout = linearize_trie(token_trie, self.device)
        self.last_layout = layout
        h = self._trunk(model, layout)
        _, W = _decoder_and_head(model)
        if self._use_fused_head(h, W):
            lp_nodes = self._head_forward_fused(h, W, layout, self.head_chunk_rows)[0]
        else:
            lp_nodes = self._head_forward(h, W, layout, self.head_chunk_rows)[0]
        for i, attach_list in enumerate(token_trie.attach_lists):
            for attach, length in attach_list:
                path = layout.path_of_leaf(i, length).long()
                returns[attach["_sequence_batch_id"]] = lp_nodes[path[1:]].clone()
        return returns

    def backward(self, model, token_trie, loss_fn: Callable, block_size: Optional[int] = 2048,
                 cut_f1_tail: bool = True) -> float:
        """Accumulate d(sum of per-sequence losses)/d(params) into ``param.grad``; return the loss
        (reference tree_training_engine.py:555-616)."""
        self.model = model
        self._check_lengths(token_trie)
        if not token_trie.inputs:
            return 0.0
        n_tok = sum(token_trie.lens) - sum(token_trie.lcp_lens)      # trie tokens (trie.py:8-13)
        budget, rec = self._plan_memory(model, n_tok, training=True)
        ranges = self._ranges(token_trie, budget) if budget is not None else None
        self.last_n_ranges = len(ranges) if ranges else 1
        if self.pass_count_listener is not None:
            self.pass_count_listener(self.last_n_ranges)
        if ranges:
            total = 0.0
            for sub in ranges:
                sub.backward_permute()
                total += self.backward_layout(model, sub, linearize_trie(sub, self.device), loss_fn, block_size, rec)
            return total
        layout = linearize_trie(token_trie, self.device)
        return self.backward_layout(model, token_trie, layout, loss_fn, block_size, rec)

    def backward_layout(self, model, token_trie, layout: TrieLayout, loss_fn: Callable,
                        block_size: Optional[int] = 2048, recompute: Optional[bool] = None) -> float:
        """``backward`` for a trie whose HBM layout is already resident (skips the H2D copy and the
        linearisation kernels; used to time the device-resident hot path)."""
        self.last_layout = layout
        if recompute is None:
            self.model = model
            recompute = self._plan_memory(model, layout.n_nodes, training=True)[1]
        self.last_recompute = bool(recompute)
        h = self._trunk(model, layout, recompute)
        _, W = _decoder_and_head(model)
        loss = _TreeHeadLoss.apply(h, W, self, layout, token_trie, loss_fn, self._chunk_rows(block_size))
        loss.backward()
        return float(loss.item())
