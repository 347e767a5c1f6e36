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

from typing import Optional, Tuple

import torch
from torch import distributed as dist

from . import _cabi

__all__ = ["gather_logprobs", "gather_logprobs_entropy", "logprob_entropy_rows"]

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


def _forward_stats(x: torch.Tensor, labels: torch.Tensor, row_of: Optional[torch.Tensor], inv_temp: float,
                   tp_group) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Shared forward: returns (x_label[M], logz[R], entropy[R]) — global over the TP group."""
    lib = _cabi.load()
    _cabi.require_cuda(x, labels)
    if x.dtype not in _DTYPE_CODE:
        raise _cabi.DtaError(f"logits dtype {x.dtype} not supported (bf16 or fp32)")
    R, V = x.shape
    M = labels.numel()
    world, rank = _tp_info(tp_group)
    dev = x.device
    st = _cabi.stream_ptr(dev)
    code = _DTYPE_CODE[x.dtype]
    labels = labels.reshape(-1).to(torch.int64)
    one_per_row = row_of is None
    if one_per_row and M != R:
        raise _cabi.DtaError(f"labels ({M}) and logits rows ({R}) disagree")
    with torch.cuda.device(dev):
        stats = torch.empty((R, 4), dtype=torch.float32, device=dev)
        _cabi.check(lib.dta_logprob_row_stats(x.data_ptr(), code, R, V, x.stride(0), inv_temp, stats.data_ptr(), st))
        if one_per_row:
            # pack x[label] into the spare 4th word: the whole exchange is 16 B per row per rank
            _cabi.check(lib.dta_logprob_gather_labels(x.data_ptr(), code, V, x.stride(0), inv_temp,
                                                      labels.data_ptr(), None, M, rank * V,
                                                      stats.data_ptr() + 12, 4, st))
            x_label = None
        else:
            x_label = torch.empty(M, dtype=torch.float32, device=dev)
            _cabi.check(lib.dta_logprob_gather_labels(x.data_ptr(), code, V, x.stride(0), inv_temp,
                                                      labels.data_ptr(), row_of.data_ptr(), M, rank * V,
                                                      x_label.data_ptr(), 1, st))
        if world > 1:
            stats_all = torch.empty((world, R, 4), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(stats_all, stats, group=tp_group)
            if x_label is not None:
                dist.all_reduce(x_label, op=dist.ReduceOp.SUM, group=tp_group)
        else:
            stats_all = stats
        logz = torch.empty(R, dtype=torch.float32, device=dev)
        entropy = torch.empty(R, dtype=torch.float32, device=dev)
        xsum = torch.empty(R, dtype=torch.float32, device=dev) if one_per_row else None
        _cabi.check(lib.dta_logprob_finalize(stats_all.data_ptr(), world, R, logz.data_ptr(), entropy.data_ptr(),
                                             xsum.data_ptr() if one_per_row else None, st))
    return (xsum if one_per_row else x_label), logz, entropy


def _backward_logits(x: torch.Tensor, labels: torch.Tensor, row_of: Optional[torch.Tensor], inv_temp: float,
                     vocab_start: int, logz: torch.Tensor, entropy: torch.Tensor, g_lp: Optional[torch.Tensor],
                     g_ent: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """d loss / d logits for this shard; ``out`` may alias ``x`` (in place)."""
    lib = _cabi.load()
    R, V = x.shape
    dev = x.device
    st = _cabi.stream_ptr(dev)
    code = _DTYPE_CODE[x.dtype]
    labels = labels.reshape(-1).to(torch.int64)
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty_like(x)
        rowsum = None
        if g_lp is not None:
            g_lp = g_lp.reshape(-1).to(torch.float32).contiguous()
            if row_of is None:
                rowsum = g_lp
            else:
                rowsum = torch.zeros(R, dtype=torch.float32, device=dev)
                _cabi.check(lib.dta_logprob_rowsum_grad(g_lp.data_ptr(), row_of.data_ptr(), g_lp.numel(),
                                                        rowsum.data_ptr(), st))
        if g_ent is not None:
            g_ent = g_ent.reshape(-1).to(torch.float32).contiguous()
        _cabi.check(lib.dta_logprob_bwd(x.data_ptr(), code, R, V, x.stride(0), inv_temp, logz.data_ptr(),
                                        entropy.data_ptr(), g_ent.data_ptr() if g_ent is not None else None,
                                        rowsum.data_ptr() if rowsum is not None else None, out.data_ptr(),
                                        out.stride(0), st))
        if g_lp is not None:
            _cabi.check(lib.dta_logprob_bwd_scatter(out.data_ptr(), code, V, out.stride(0), inv_temp,
                                                    labels.data_ptr(),
                                                    row_of.data_ptr() if row_of is not None else None,
                                                    g_lp.data_ptr(), g_lp.numel(), vocab_start, st))
    return out


class _LogProbEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, inv_temp: float, tp_group, want_entropy: bool):
        x = _as_rows(logits)
        x_label, logz, entropy = _forward_stats(x, labels, None, inv_temp, tp_group)
        world, rank = _tp_info(tp_group)
        ctx.save_for_backward(x, labels, logz, entropy)
        ctx.inv_temp = inv_temp
        ctx.vocab_start = rank * x.shape[1]
        ctx.in_shape = logits.shape
        ctx.want_entropy = want_entropy
        lp = (x_label - logz).reshape(labels.shape)
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
    return _forward_stats(x, labels, row_of, inv_temp, tp_group)


logprob_entropy_rows_backward = _backward_logits
