"""Vocab-parallel log-prob / entropy HOST logic on CPU: world_size-2/-4 gloo (BASELINE config 4).

The ``tp_group`` branch of ``dynamictreeattn_b200.vocab_parallel`` (shard statistics block -> ONE all-gather ->
online-softmax merge -> collective-free backward) runs here with the six kernel wrappers replaced by the torch
restatements of tests/fake_backend.py.  The reference's TP branch has no caller and no recorded result; its only
oracle is algebraic equality with the non-TP functions on the concatenated logits (vocab_parallel.py:13-27), which
is what these tests check, for one label per row (the public ``gather_*`` API) and for several labels per row
(``row_of``: trie edges)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _problem(R=23, V=96, M=40, seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(R, V, generator=g) * 3
    labels = torch.randint(0, V, (R,), generator=g)
    w_lp, w_ent = torch.randn(R, generator=g), torch.randn(R, generator=g)
    row_of = torch.randint(0, R, (M,), generator=g).to(torch.int32)
    row_of[3] = -1                                          # a root node: no edge
    edge_labels = torch.randint(0, V, (M,), generator=g)
    g_edge = torch.randn(M, generator=g)
    g_edge[3] = 0
    return logits, labels, w_lp, w_ent, row_of, edge_labels, g_edge


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_backend
    fake_backend.install_vocab_parallel(_MP())
    from dynamictreeattn_b200 import vocab_parallel as vp
    calls = {"all_gather": 0, "all_reduce": 0}
    real_ag, real_ar = dist.all_gather_into_tensor, dist.all_reduce

    def count_ag(*a, **k):
        calls["all_gather"] += 1
        return real_ag(*a, **k)

    def count_ar(*a, **k):
        calls["all_reduce"] += 1
        return real_ar(*a, **k)
    dist.all_gather_into_tensor, dist.all_reduce = count_ag, count_ar

    logits, labels, w_lp, w_ent, row_of, edge_labels, g_edge = _problem()
    V = logits.shape[1]
    Vs = V // world
    shard = logits[:, rank * Vs:(rank + 1) * Vs].contiguous().requires_grad_(True)
    grp = dist.group.WORLD
    # public API, 3-D leading dims like the reference's callers ([1, B, V], tree_training_engine.py:189-193)
    lp, ent = vp.gather_logprobs_entropy(shard.unsqueeze(0), labels.unsqueeze(0), temperature=0.8, tp_group=grp)
    ((lp[0] * w_lp).sum() + (ent[0] * w_ent).sum()).backward()
    lp_only = vp.gather_logprobs(shard.detach(), labels, temperature=0.8, tp_group=grp)
    n_collectives_public = dict(calls)
    # engine-facing primitives: several labels per row
    x_label, logz, ent_rows = vp.logprob_entropy_rows(shard.detach(), edge_labels, row_of, 1.0, grp)
    d_edge = vp.logprob_entropy_rows_backward(shard.detach(), edge_labels, row_of, 1.0, rank * Vs, logz, ent_rows,
                                              g_edge, w_ent)
    torch.save({"lp": lp.detach(), "ent": ent.detach(), "lp_only": lp_only, "grad": shard.grad, "x_label": x_label,
                "logz": logz, "ent_rows": ent_rows, "d_edge": d_edge, "calls_public": n_collectives_public,
                "calls_total": dict(calls)}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_tp_group_branch_equals_unsharded(tmp_path, world):
    from oracle import tree_oracle
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    logits, labels, w_lp, w_ent, row_of, edge_labels, g_edge = _problem()
    V = logits.shape[1]
    Vs = V // world
    # ---- oracle: the non-TP definition on the full logits (vocab_parallel.py:13-27) -----------------------------
    ref = logits.clone().requires_grad_(True)
    lpr, entr = tree_oracle.logprob_entropy(ref, labels, 0.8)
    ((lpr * w_lp).sum() + (entr * w_ent).sum()).backward()
    for r, o in enumerate(outs):
        assert o["lp"].shape == (1, logits.shape[0]) and o["lp"].dtype == torch.float32
        assert (o["lp"][0] - lpr.detach()).abs().max().item() < 1e-5          # every rank holds the GLOBAL result
        assert (o["ent"][0] - entr.detach()).abs().max().item() < 1e-5
        assert (o["lp_only"] - lpr.detach()).abs().max().item() < 1e-5
        gs = ref.grad[:, r * Vs:(r + 1) * Vs]
        assert ((o["grad"] - gs).norm() / gs.norm()).item() < 1e-5            # backward: own shard, no collective
        # ONE all-gather per forward, nothing else (reference: 3 / 4 all-reduces); backward collective-free
        assert o["calls_public"] == {"all_gather": 2, "all_reduce": 0}
        assert o["calls_total"] == {"all_gather": 3, "all_reduce": 0}
    # ---- several labels per row --------------------------------------------------------------------------------
    ref2 = logits.clone().requires_grad_(True)
    lsm = torch.log_softmax(ref2, -1)
    valid = row_of >= 0
    ent2 = -(lsm.exp() * lsm).sum(-1)
    ((lsm[row_of[valid].long(), edge_labels[valid]] * g_edge[valid]).sum() + (ent2 * w_ent).sum()).backward()
    want_x = torch.where(valid, logits[row_of.clamp(min=0).long(), edge_labels], torch.zeros(()))
    for r, o in enumerate(outs):
        assert (o["x_label"] - want_x).abs().max().item() < 1e-6
        assert (o["logz"] - torch.logsumexp(logits, -1)).abs().max().item() < 1e-5
        assert (o["ent_rows"] - ent2.detach()).abs().max().item() < 1e-5
        gs = ref2.grad[:, r * Vs:(r + 1) * Vs]
        assert ((o["d_edge"] - gs).norm() / gs.norm()).item() < 1e-5


def test_single_rank_and_error_contracts(monkeypatch):
    """world == 1 (no group / group of one) takes the same code without a collective; shape errors as the reference
    would raise them."""
    import fake_backend
    fake_backend.install_vocab_parallel(monkeypatch)
    from dynamictreeattn_b200 import vocab_parallel as vp
    from oracle import tree_oracle
    logits, labels, w_lp, w_ent, *_ = _problem(R=7, V=33)
    x = logits.clone().requires_grad_(True)
    lp, ent = vp.gather_logprobs_entropy(x, labels, temperature=1.3)
    ((lp * w_lp).sum() + (ent * w_ent).sum()).backward()
    ref = logits.clone().requires_grad_(True)
    lpr, entr = tree_oracle.logprob_entropy(ref, labels, 1.3)
    ((lpr * w_lp).sum() + (entr * w_ent).sum()).backward()
    assert (lp - lpr).abs().max().item() < 1e-5 and (ent - entr).abs().max().item() < 1e-5
    assert ((x.grad - ref.grad).norm() / ref.grad.norm()).item() < 1e-5
    with pytest.raises(ValueError):
        vp.gather_logprobs(logits, labels[:-1])
