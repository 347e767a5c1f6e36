"""Seeded synthetic trie workloads: the tau2-16k-shaped trie of BASELINE config 2, the kernel-sweep
families of config 5 and small adversarial mixes.  Used by bench.py, the tools and the tests; the reference ships no
data (its ``data/tau2-16k-merged`` dump is a stripped large blob), so every measurement runs on these."""
from __future__ import annotations

import random
from typing import List

import torch


def random_seqs(seed: int, n_seq: int, vocab: int = 5, max_len: int = 12) -> List[torch.Tensor]:
    """Small adversarial mixes: shared prefixes, full-prefix sequences, duplicates, forests."""
    rng = random.Random(seed)
    seqs: List[List[int]] = []
    for _ in range(n_seq):
        if seqs and rng.random() < 0.7:
            p = rng.choice(seqs)
            cut = rng.randrange(0, len(p) + 1)
            s = p[:cut] + [rng.randrange(vocab) for _ in range(rng.randrange(0, max_len - cut + 1))]
        else:
            s = [rng.randrange(vocab) for _ in range(rng.randrange(1, max_len))]
        seqs.append(s or [0])
    return [torch.tensor(s, dtype=torch.long) for s in seqs]


def synth_trie(seed: int, trunk: int, fanout=(4, 2, 2), seg=(1024, 3072), vocab: int = 151936,
               max_depth: int = 16384) -> List[torch.Tensor]:
    """tau2-shaped rollouts: a shared trunk, then fork levels with the given
    fan-outs; every segment length ~ U{seg}; siblings start with distinct tokens; paths are
    truncated to max_depth.  Returns one sequence per leaf (shuffled deterministically)."""
    g = torch.Generator().manual_seed(seed)

    def rand_ids(n):
        return torch.randint(0, vocab, (n,), generator=g)

    def seg_len():
        return int(torch.randint(seg[0], seg[1] + 1, (1,), generator=g))

    paths = [rand_ids(trunk)]
    for fan in fanout:
        nxt = []
        for p in paths:
            firsts = torch.randperm(vocab, generator=g)[:fan]
            for c in range(fan):
                s = rand_ids(seg_len())
                s[0] = firsts[c]
                nxt.append(torch.cat([p, s])[:max_depth])
        paths = nxt
    order = torch.randperm(len(paths), generator=g).tolist()
    return [paths[i] for i in order]


def config1_seqs() -> List[torch.Tensor]:
    """BASELINE config 1 trie: trunk 64, 4 mids x 32, 2 tails (16, 24) + the 40-token pure prefix."""
    g = torch.Generator().manual_seed(1)
    trunk = torch.randint(0, 512, (64,), generator=g)
    seqs = []
    for m in range(4):
        mid = torch.randint(0, 512, (32,), generator=g)
        mid[0] = m
        for t, tl in enumerate((16, 24)):
            tail = torch.randint(0, 512, (tl,), generator=g)
            tail[0] = t
            seqs.append(torch.cat([trunk, mid, tail]))
    seqs.append(trunk[:40].clone())
    return seqs


def trunk_fan(n_total: int, trunk: int, fan: int, vocab: int = 50000, seed: int = 0) -> List[torch.Tensor]:
    """BASELINE config 5, family A: a shared trunk of ``trunk`` tokens, then ``fan`` equal leaves."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(0, vocab, (trunk,), generator=g)
    leaf = (n_total - trunk) // fan
    firsts = torch.randperm(vocab, generator=g)[:fan]
    out = []
    for c in range(fan):
        s = torch.randint(0, vocab, (leaf,), generator=g)
        s[0] = firsts[c]
        out.append(torch.cat([t, s]))
    return out


def bary(n_total: int, fan: int, levels: int, vocab: int = 50000, seed: int = 0) -> List[torch.Tensor]:
    """BASELINE config 5, family B: fan-out ``fan`` at each of ``levels`` levels, equal segments."""
    n_seg = sum(fan ** l for l in range(levels + 1))
    seg = max(n_total // n_seg, 1)
    return synth_trie(seed, seg, tuple([fan] * levels), (seg, seg), vocab=vocab, max_depth=1 << 30)


def tau2_trie(seed: int = 0) -> List[torch.Tensor]:
    """BASELINE config 2 / SURVEY.md §8d: trunk 6144, fan-out (4,2,2), segments U{1024..3072}, depth <= 16384
    (seed 0: 16 sequences, 67,183 trie tokens, 202,574 sequence tokens, 6.52e8 visible pairs)."""
    return synth_trie(seed, 6144, (4, 2, 2), (1024, 3072), vocab=151936, max_depth=16384)


def cpu_sample_trie(seed: int = 0) -> List[torch.Tensor]:
    """Bounded sample of the tau2 generator for CPU legs and full-model parity checks (~2.4k trie tokens)."""
    return synth_trie(seed, 512, (2, 2), (128, 384), vocab=151936, max_depth=16384)
