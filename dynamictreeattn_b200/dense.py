"""Per-sequence dense-causal baseline — drop-in for the reference module ``dense.py`` (dense.py:7-53).

One HF causal forward (and backward) per sequence, no prefix sharing.  It is the second parity
oracle of the tree path and the GPU baseline the reference compares against (run.py:11-26, 63-77).
Log-probs / entropies come from the fused sm_100a kernel (``vocab_parallel.py``), so on a GPU this is
"stock HF attention + our logprob op"; the attention backend is whatever the model was loaded with.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch

from .vocab_parallel import gather_logprobs, gather_logprobs_entropy

__all__ = ["forward", "backward"]


def _iter(n: int, use_tqdm: bool):
    if use_tqdm:
        import tqdm
        return tqdm.tqdm(range(n))
    return range(n)


@torch.no_grad()
def forward(model, token_seqs: Sequence[torch.Tensor], use_tqdm: bool = False) -> List[torch.Tensor]:
    out = []
    for i in _iter(len(token_seqs), use_tqdm):
        ids = token_seqs[i].unsqueeze(0).to(model.device)
        logits = model(input_ids=ids, use_cache=False).logits
        out.append(gather_logprobs(logits=logits[:, :-1], labels=ids[:, 1:]).squeeze(0))
    return out


def backward(model, token_seqs: Sequence[torch.Tensor], attachs: Sequence[dict], loss_fn: Callable,
             act_ckpt: bool = False, use_tqdm: bool = False) -> float:
    if act_ckpt:
        model.gradient_checkpointing_enable()
    else:
        model.gradient_checkpointing_disable()
    total = 0.0
    for i in _iter(len(token_seqs), use_tqdm):
        ids = token_seqs[i].unsqueeze(0).to(model.device)
        logits = model(input_ids=ids, use_cache=False).logits
        # entropy has length L, log-probs L-1 (dense.py:41-44): the last row's label is a dummy
        labels = torch.cat([ids[:, 1:], ids[:, :1]], dim=1)
        lp, ent = gather_logprobs_entropy(logits=logits, labels=labels)
        loss = loss_fn(lp.squeeze(0)[:-1], ent.squeeze(0), attachs[i])
        loss.backward()
        total += float(loss.item())
    return total
