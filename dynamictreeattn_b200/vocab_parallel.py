"""Per-token log-prob and entropy over the vocabulary, optionally vocab-sharded.

Drop-in for the reference module ``vocab_parallel.py`` (vocab_parallel.py:399-467):
``gather_logprobs`` / ``gather_logprobs_entropy`` keep their signatures, accept
arbitrary leading dims, return fp32 and are differentiable w.r.t. ``logits``.

What runs is different:

* one fused sm_100a kernel reads each logit ONCE and emits 16 bytes per row
  ``{max, sum e^(x-max), sum x e^(x-max), x[label]}`` — no fp32 ``[rows, vocab]``
  log-softmax tensor is ever materialised (reference: vocab_parallel.py:13-27);
* with ``tp_group`` the ranks exchange exactly those 16 bytes per row in ONE
  ``all_gather`` and merge them with the online-softmax rule, instead of the
  reference's 3-4 latency-bound all-reduces (vocab_parallel.py:134,142,156 /
  264,273,291,298);
* the backward needs no saved softmax (reference saves fp32 ``[rows, vocab/tp]``,
  vocab_parallel.py:312-318): it re-reads the logits once and writes the gradient once.

The same kernels serve the tree engine, where one logits row can carry several labels
(one per child edge); that case is expressed with ``row_of`` (label -> row).
"""
from __future__ import annotations

from contextlib import nullcontext as _nullcontext
from typing import Optional, Tuple

import torch
from torch import distributed as dist

from . import _cabi

__all__ = ["gather_logprobs", "gather_logprobs_entropy", "logprob_entropy_rows", "lm_head_logprob_entropy_rows",
           "lm_head_dlogits_rows", "lm_head_logits_and_stats_rows", "fused_head_supported"]

_DTYPE_CODE = {torch.bfloat16: 0, torch.float32: 1}


def _as_rows(logits: torch.Tensor) -> torch.Tensor:
    """[..., V] -> [R, V] view with unit column stride (copy only if unavoidable)."""
    x = logits.reshape(-1, logits.shape[-1])
    if x.stride(-1) != 1:
        x = x.contiguous()
    return x


def _tp_info(tp_group) -> Tuple[int, int]:
    if tp_group is not None and dist.is_initialized():
        world = dist.get_world_size(tp_group)
        if world > 1:
            return world, dist.get_rank(tp_group)
    return 1, 0


# ---------------------------------------------------------------------------------------------------------
# kernel wrappers: the ONLY functions of this module that touch the C-ABI.  Everything above them (sharding,
# the collective, autograd) is plain host logic; the CPU test-suite swaps these six for torch restatements
# (tests/fake_backend.py) to run that logic under gloo.
# ---------------------------------------------------------------------------------------------------------
def _k_row_stats(x: torch.Tensor, inv_temp: float, labels: Optional[torch.Tensor], vocab_start: int,
                 stats: torch.Tensor) -> None:
    """stats[R,4] <- {max, sum e^(x-max), sum x e^(x-max), x[label] (one label per row, 0 outside the shard)}."""
    lib = _cabi.load()
    R, V = x.shape
    _cabi.check(lib.dta_logprob_row_stats(x.data_ptr(), _DTYPE_CODE[x.dtype], R, V, x.stride(0), inv_temp,
                                          labels.data_ptr() if labels is not None else None, vocab_start,
                                          stats.data_ptr(), _cabi.stream_ptr(x.device)))


def _k_gather_labels(x: torch.Tensor, inv_temp: float, labels: torch.Tensor, row_of: torch.Tensor, vocab_start: int,
                     out: torch.Tensor) -> None:
    """out[m] <- x[row_of[m]][labels[m] - vocab_start] for labels in the shard, else 0 (several labels per row)."""
    lib = _cabi.load()
    _cabi.check(lib.dta_logprob_gather_labels(x.data_ptr(), _DTYPE_CODE[x.dtype], x.shape[1], x.stride(0), inv_temp,
                                              labels.data_ptr(), row_of.data_ptr(), labels.numel(), vocab_start,
                                              out.data_ptr(), 1, _cabi.stream_ptr(x.device)))


def _k_finalize(blocks: torch.Tensor, n_rows: int, want_lp: bool):
    """blocks [world, >= 4 n_rows] fp32 (row statistics first) -> (logz, entropy, logprob or None), each [n_rows]."""
    lib = _cabi.load()
    dev = blocks.device
    logz = torch.empty(n_rows, dtype=torch.float32, device=dev)
    entropy = torch.empty(n_rows, dtype=torch.float32, device=dev)
    lp = torch.empty(n_rows, dtype=torch.float32, device=dev) if want_lp else None
    _cabi.check(lib.dta_logprob_finalize(blocks.data_ptr(), blocks.shape[0], blocks.stride(0), n_rows, logz.data_ptr(),
                                         entropy.data_ptr(), lp.data_ptr() if want_lp else None,
                                         _cabi.stream_ptr(dev)))
    return logz, entropy, lp


def _k_backward(x: torch.Tensor, inv_temp: float, logz: torch.Tensor, entropy: torch.Tensor,
                g_ent: Optional[torch.Tensor], g_rowsum: Optional[torch.Tensor], row_labels: Optional[torch.Tensor],
                vocab_start: int, out: torch.Tensor) -> None:
    """out <- d loss / d logits; with ``row_labels`` (one per row) the one-hot term is added in the same pass."""
    lib = _cabi.load()
    R, V = x.shape
    _cabi.check(lib.dta_logprob_bwd(x.data_ptr(), _DTYPE_CODE[x.dtype], R, V, x.stride(0), inv_temp, logz.data_ptr(),
                                    entropy.data_ptr(), g_ent.data_ptr() if g_ent is not None else None,
                                    g_rowsum.data_ptr() if g_rowsum is not None else None,
                                    row_labels.data_ptr() if row_labels is not None else None, vocab_start,
                                    out.data_ptr(), out.stride(0), _cabi.stream_ptr(x.device)))


def _k_rowsum(g_lp: torch.Tensor, row_of: torch.Tensor, n_rows: int) -> torch.Tensor:
    """rowsum[r] = sum of g_lp over the labels attached to row r."""
    lib = _cabi.load()
    out = torch.zeros(n_rows, dtype=torch.float32, device=g_lp.device)
    _cabi.check(lib.dta_logprob_rowsum_grad(g_lp.data_ptr(), row_of.data_ptr(), g_lp.numel(), out.data_ptr(),
                                            _cabi.stream_ptr(g_lp.device)))
    return out


def _k_scatter(out: torch.Tensor, inv_temp: float, labels: torch.Tensor, row_of: torch.Tensor, g_lp: torch.Tensor,
               vocab_start: int) -> None:
    """out[row_of[m]][labels[m] - vocab_start] += inv_temp * g_lp[m] for labels in the shard (several per row)."""
    lib = _cabi.load()
    _cabi.check(lib.dta_logprob_bwd_scatter(out.data_ptr(), _DTYPE_CODE[out.dtype], out.shape[1], out.stride(0),
                                            inv_temp, labels.data_ptr(), row_of.data_ptr(), g_lp.data_ptr(),
                                            g_lp.numel(), vocab_start, _cabi.stream_ptr(out.device)))


def _forward_stats(x: torch.Tensor, labels: torch.Tensor, row_of: Optional[torch.Tensor], inv_temp: float,
                   tp_group) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], torch.Tensor, torch.Tensor]:
    """Shared forward -> (logprob[R] | None, x_label[M] | None, logz[R], entropy[R]), global over the TP group.

    Each rank fills ONE block of fp32 words — 16 bytes of statistics per row (the label logit rides in the 4th
    word), followed, for trie rows with several labels (``row_of``), by one word per label — and the group
    exchanges exactly that block in ONE all-gather (reference: 3-4 all-reduces, vocab_parallel.py:134,142,156 /
    264,273,291,298)."""
    _cabi.require_cuda(x, labels)
    if x.dtype not in _DTYPE_CODE:
        raise _cabi.DtaError(f"logits dtype {x.dtype} not supported (bf16 or fp32)")
    R, V = x.shape
    M = labels.numel()
    world, rank = _tp_info(tp_group)
    dev = x.device
    labels = labels.reshape(-1).to(torch.int64)
    one_per_row = row_of is None
    if one_per_row and M != R:
        raise _cabi.DtaError(f"labels ({M}) and logits rows ({R}) disagree")
    n_extra = 0 if one_per_row else (M + 3) // 4 * 4          # blocks stay 16-byte aligned
    with torch.cuda.device(dev) if dev.type == "cuda" else _nullcontext():
        block = torch.empty(4 * R + n_extra, dtype=torch.float32, device=dev)
        _k_row_stats(x, inv_temp, labels if one_per_row else None, rank * V, block[:4 * R].view(R, 4))
        if not one_per_row:
            if n_extra > M:
                block[4 * R + M:].zero_()
            _k_gather_labels(x, inv_temp, labels, row_of, rank * V, block[4 * R:4 * R + M])
        if world > 1:
            blocks = torch.empty((world, block.numel()), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(blocks.view(-1), block, group=tp_group)   # flat output: gloo and NCCL both take it
        else:
            blocks = block.view(1, -1)
        logz, entropy, lp = _k_finalize(blocks, R, one_per_row)
        x_label = None
        if not one_per_row:
            x_label = blocks[:, 4 * R:4 * R + M].sum(0) if world > 1 else block[4 * R:4 * R + M]
    return lp, x_label, logz, entropy


def _backward_logits(x: torch.Tensor, labels: torch.Tensor, row_of: Optional[torch.Tensor], inv_temp: float,
                     vocab_start: int, logz: torch.Tensor, entropy: torch.Tensor, g_lp: Optional[torch.Tensor],
                     g_ent: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """d loss / d logits for this shard (collective-free, like the reference's backward, vocab_parallel.py:324-370);
    ``out`` may alias ``x`` (in place)."""
    R, V = x.shape
    dev = x.device
    labels = labels.reshape(-1).to(torch.int64)
    with torch.cuda.device(dev) if dev.type == "cuda" else _nullcontext():
        if out is None:
            out = torch.empty_like(x)
        if g_lp is not None:
            g_lp = g_lp.reshape(-1).to(torch.float32).contiguous()
        if g_ent is not None:
            g_ent = g_ent.reshape(-1).to(torch.float32).contiguous()
        if row_of is None:        # one label per row: a single launch
            _k_backward(x, inv_temp, logz, entropy, g_ent, g_lp, labels if g_lp is not None else None, vocab_start, out)
        else:
            rowsum = _k_rowsum(g_lp, row_of, R) if g_lp is not None else None
            _k_backward(x, inv_temp, logz, entropy, g_ent, rowsum, None, vocab_start, out)
            if g_lp is not None:
                _k_scatter(out, inv_temp, labels, row_of, g_lp, vocab_start)
    return out


class _LogProbEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, inv_temp: float, tp_group, want_entropy: bool):
        x = _as_rows(logits)
        lp, _, logz, entropy = _forward_stats(x, labels, None, inv_temp, tp_group)
        world, rank = _tp_info(tp_group)
        ctx.save_for_backward(x, labels, logz, entropy)
        ctx.inv_temp = inv_temp
        ctx.vocab_start = rank * x.shape[1]
        ctx.in_shape = logits.shape
        ctx.want_entropy = want_entropy
        lp = lp.reshape(labels.shape)
        if want_entropy:
            return lp, entropy.reshape(labels.shape)
        return lp

    @staticmethod
    def backward(ctx, g_lp, g_ent=None):
        x, labels, logz, entropy = ctx.saved_tensors
        grad = _backward_logits(x, labels, None, ctx.inv_temp, ctx.vocab_start, logz, entropy, g_lp,
                                g_ent if ctx.want_entropy else None)
        return grad.reshape(ctx.in_shape), None, None, None, None


def _prep(logits: torch.Tensor, labels: torch.Tensor):
    if labels.shape != logits.shape[:-1]:
        raise ValueError(f"labels shape {tuple(labels.shape)} must equal logits leading dims {tuple(logits.shape[:-1])}")
    if logits.dtype not in _DTYPE_CODE:
        logits = logits.float()
    return logits, labels


def gather_logprobs(logits: torch.Tensor, labels: torch.Tensor, temperature: float = 1.0,
                    tp_group: Optional[dist.ProcessGroup] = None, chunk_size: int = 1024) -> torch.Tensor:
    """log softmax(logits / temperature)[labels], fp32 (reference vocab_parallel.py:399-432).

    ``chunk_size`` is accepted for signature compatibility; the fused kernel streams
    the logits row by row, so there is nothing to chunk.
    """
    logits, labels = _prep(logits, labels)
    return _LogProbEntropy.apply(logits, labels, 1.0 / float(temperature), tp_group, False)


def gather_logprobs_entropy(logits: torch.Tensor, labels: torch.Tensor, temperature: float = 1.0,
                            tp_group: Optional[dist.ProcessGroup] = None,
                            chunk_size: int = 1024) -> Tuple[torch.Tensor, torch.Tensor]:
    """(log-prob at labels, entropy of softmax(logits / temperature)), both fp32
    (reference vocab_parallel.py:435-467)."""
    logits, labels = _prep(logits, labels)
    return _LogProbEntropy.apply(logits, labels, 1.0 / float(temperature), tp_group, True)


# engine-facing primitives (several labels per row through ``row_of``)
def logprob_entropy_rows(x: torch.Tensor, labels: torch.Tensor, row_of: torch.Tensor, inv_temp: float = 1.0,
                         tp_group=None):
    """(x_label[M], logz[R], entropy[R]) for logits rows ``x`` [R, V]; no autograd."""
    _, x_label, logz, entropy = _forward_stats(x, labels, row_of, inv_temp, tp_group)
    return x_label, logz, entropy


logprob_entropy_rows_backward = _backward_logits


# ---------------------------------------------------------------------------------------------------------
# N1: lm_head fused into the statistics -- the [rows, vocab] logits are never materialised (csrc/head_fwd.cu)
# ---------------------------------------------------------------------------------------------------------
def lm_head_logprob_entropy_rows(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                                 row_of: Optional[torch.Tensor], inv_temp: float = 1.0):
    """(x_label[M], logz[R], entropy[R]) of ``logits = hidden @ weight.T`` WITHOUT the logits: a tcgen05 GEMM whose
    accumulator tiles are folded into the 16-byte row statistics in the epilogue (reference: HF
    modeling_qwen3.py:505 + vocab_parallel.py:13-27 over a materialised [B, V] tensor).  hidden [R, K], weight [V, K]
    bf16 with unit inner stride; ``row_of`` (label -> row; None: one label per row) as in ``logprob_entropy_rows``.
    The logits behind these numbers are the fp32 accumulators, not their bf16 roundings."""
    lib = _cabi.load()
    _cabi.require_cuda(hidden, weight, labels)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise _cabi.DtaError("the fused lm_head kernel takes bf16 hidden states and weights")
    if hidden.dim() != 2 or weight.dim() != 2 or hidden.shape[1] != weight.shape[1]:
        raise _cabi.DtaError(f"hidden {tuple(hidden.shape)} and weight {tuple(weight.shape)} do not form a GEMM")
    if hidden.stride(1) != 1:
        hidden = hidden.contiguous()
    if weight.stride(1) != 1:
        weight = weight.contiguous()
    R, K = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    labels = labels.reshape(-1).to(torch.int64)
    M = labels.numel()
    if row_of is None and M != R:
        raise _cabi.DtaError(f"labels ({M}) and rows ({R}) disagree")
    st = _cabi.stream_ptr(dev)
    with torch.cuda.device(dev):
        n_split = lib.dta_lm_head_stats_splits(R, V)
        rows_pad = lib.dta_lm_head_stats_rows_pad(R)
        stats = torch.empty((n_split, rows_pad, 4), dtype=torch.float32, device=dev)
        _cabi.check(lib.dta_lm_head_stats(hidden.data_ptr(), hidden.stride(0), weight.data_ptr(), weight.stride(0), R, V,
                                          K, inv_temp, stats.data_ptr(), st))
        logz, entropy, _ = _k_finalize(stats.view(n_split, rows_pad * 4), R, False)
        x_label = torch.empty(M, dtype=torch.float32, device=dev)
        _cabi.check(lib.dta_lm_head_label_logits(hidden.data_ptr(), hidden.stride(0), weight.data_ptr(), weight.stride(0),
                                                 K, V, labels.data_ptr(),
                                                 row_of.data_ptr() if row_of is not None else None, M, inv_temp,
                                                 x_label.data_ptr(), st))
    return x_label, logz, entropy


def lm_head_logits_and_stats_rows(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                                  row_of: torch.Tensor, inv_temp: float = 1.0):
    """(logits [R, V] bf16, x_label[M], logz[R], entropy[R]): the fused lm_head launch that ALSO keeps the logits for a
    backward over them (``logprob_entropy_rows_backward``).  Statistics and label logits are those of the stored bf16
    tensor -- the same numbers a library GEMM followed by ``logprob_entropy_rows`` gives, without reading the logits
    back in the forward."""
    lib = _cabi.load()
    R, K = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    labels = labels.reshape(-1).to(torch.int64)
    M = labels.numel()
    st = _cabi.stream_ptr(dev)
    with torch.cuda.device(dev):
        n_split = lib.dta_lm_head_stats_splits(R, V)
        rows_pad = lib.dta_lm_head_stats_rows_pad(R)
        stats = torch.empty((n_split, rows_pad, 4), dtype=torch.float32, device=dev)
        logits = torch.empty((R, V), dtype=torch.bfloat16, device=dev)
        _cabi.check(lib.dta_lm_head_stats_logits(hidden.data_ptr(), hidden.stride(0), weight.data_ptr(), weight.stride(0), R,
                                                 V, K, inv_temp, stats.data_ptr(), logits.data_ptr(), logits.stride(0), st))
        logz, entropy, _ = _k_finalize(stats.view(n_split, rows_pad * 4), R, False)
        x_label = torch.empty(M, dtype=torch.float32, device=dev)
        _k_gather_labels(logits, inv_temp, labels, row_of, 0, x_label)      # stored logits are raw: x = logit * inv_temp
    return logits, x_label, logz, entropy


def fused_head_supported(hidden: torch.Tensor, weight: torch.Tensor) -> bool:
    """bf16 CUDA operands the TMA descriptors of csrc/head_fwd.cu can address."""
    return (hidden.is_cuda and weight.is_cuda and hidden.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and hidden.dim() == 2 and weight.dim() == 2 and hidden.shape[1] == weight.shape[1]
            and hidden.shape[1] % 64 == 0 and hidden.stride(1) == 1 and weight.stride(1) == 1
            and hidden.stride(0) % 8 == 0 and weight.stride(0) % 8 == 0
            and hidden.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0)


def lm_head_dlogits_rows(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, row_of: torch.Tensor,
                         inv_temp: float, logz: torch.Tensor, entropy: torch.Tensor, g_lp: Optional[torch.Tensor],
                         g_ent: Optional[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
    """d loss / d logits [R, V] (bf16, into ``out``) for ``logits = hidden @ weight.T`` WITHOUT kept logits: the GEMM is
    recomputed by the fused kernel whose epilogue applies the log-prob / entropy backward; the one-hot terms of the
    labels (``row_of``: label -> row of this block) are added afterwards.  Counterpart of
    ``logprob_entropy_rows_backward`` for the fused head."""
    lib = _cabi.load()
    R, K = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    st = _cabi.stream_ptr(dev)
    labels = labels.reshape(-1).to(torch.int64)
    with torch.cuda.device(dev):
        rowsum = None
        if g_lp is not None:
            g_lp = g_lp.reshape(-1).to(torch.float32).contiguous()
            rowsum = _k_rowsum(g_lp, row_of, R)
        if g_ent is not None:
            g_ent = g_ent.reshape(-1).to(torch.float32).contiguous()
        _cabi.check(lib.dta_lm_head_dlogits(hidden.data_ptr(), hidden.stride(0), weight.data_ptr(), weight.stride(0), R, V, K,
                                            inv_temp, logz.data_ptr(), entropy.data_ptr(),
                                            g_ent.data_ptr() if g_ent is not None else None,
                                            rowsum.data_ptr() if rowsum is not None else None, out.data_ptr(),
                                            out.stride(0), st))
        if g_lp is not None:
            _k_scatter(out, inv_temp, labels, row_of, g_lp, 0)
    return out
