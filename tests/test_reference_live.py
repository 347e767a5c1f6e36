"""Live differential tests against the reference checkout (build container only; skipped on the GPU box)."""
import random
from types import SimpleNamespace

import pytest
import torch

from dynamictreeattn_b200 import data_parallel as dp
from dynamictreeattn_b200 import token_trie as mtt
from dynamictreeattn_b200 import tree_time_model as mtm
from dynamictreeattn_b200 import trie as mtrie
from oracle import tree_oracle
from trie_cases import config1_seqs, random_seqs

pytestmark = pytest.mark.reference


def test_host_modules_differential(reference):
    rng = random.Random(0)
    for it in range(120):
        seqs = random_seqs(1000 + it, rng.randrange(1, 14))
        R = reference.token_trie.TokenTrie(seqs, [{} for _ in seqs])
        M = mtt.TokenTrie(seqs, [{} for _ in seqs])
        assert [t.tolist() for t in R.inputs] == [t.tolist() for t in M.inputs]
        assert R.lcp_lens == M.lcp_lens and R.lens == M.lens
        rc = reference.trie.CompressedTrie(R.lens, R.lcp_lens)
        mc = mtrie.CompressedTrie(M.lens, M.lcp_lens)
        assert rc.get_order_forward() == mc.get_order_forward()
        assert rc.get_order_backward() == mc.get_order_backward()
        random.seed(it); ro = rc.get_order_random()
        random.seed(it); mo = mc.get_order_random()
        assert ro == mo
        for K in (2, 4):
            args = SimpleNamespace(K=K, mode="forward", block_size=None)
            assert reference.data_parallel.LB_by_DFS_and_TM(seqs, reference.tree_time_model.TreeTimeModel(), args) == \
                dp.LB_by_DFS_and_TM(seqs, mtm.TreeTimeModel(), args)
        R.random_permute
        R.forward_permute(); M.forward_permute()
        assert R.lcp_lens == M.lcp_lens and [t.tolist() for t in R.inputs] == [t.tolist() for t in M.inputs]


def test_backward_mode_partition_on_long_sequences(reference):
    """mode='backward' with block_size (the reference's production setting) on sequences long enough
    not to trip the reference's own IndexError (trie.py:35)."""
    g = torch.Generator().manual_seed(0)
    base = torch.randint(0, 50, (40,), generator=g)
    seqs = [torch.cat([base[: 10 + 3 * i], torch.randint(50, 99, (30 + i,), generator=g)]) for i in range(9)]
    args = SimpleNamespace(K=3, mode="backward", block_size=8)
    assert reference.data_parallel.LB_by_DFS_and_TM(seqs, reference.tree_time_model.TreeTimeModel(), args) == \
        dp.LB_by_DFS_and_TM(seqs, mtm.TreeTimeModel(), args)
    assert reference.data_parallel.LB_by_TM(seqs, reference.tree_time_model.TreeTimeModel(), args) == \
        dp.LB_by_TM(seqs, mtm.TreeTimeModel(), args)


def test_oracle_stack_walk_equals_reference_engine(reference):
    from test_engine_cpu import attachs, grads, tiny_model, worst_ratio
    cfg, model = tiny_model()
    seqs = config1_seqs() + [config1_seqs()[0].clone(), torch.randint(0, 512, (30,))]
    model.zero_grad()
    eng = reference.tree_training_engine.TreeTrainingEngine(cfg, "cpu", torch.float32, 4096)
    t = reference.token_trie.TokenTrie(seqs, attachs(len(seqs)))
    t.backward_permute()
    loss_ref = eng.backward(model, t, tree_oracle.example_loss, block_size=16)
    g_ref = grads(model)
    model.zero_grad()
    t2 = mtt.TokenTrie(seqs, attachs(len(seqs)))
    t2.backward_permute()
    loss_o = tree_oracle.tree_stack_backward(model, t2.inputs, t2.lcp_lens, t2.attach_lists, tree_oracle.example_loss)
    assert abs(loss_o - loss_ref) < 1e-4
    assert worst_ratio(grads(model), g_ref) < 1e-5
    ft = reference.token_trie.TokenTrie(seqs)
    ft.forward_permute()
    feng = reference.tree_training_engine.TreeTrainingEngine(cfg, "cpu", torch.float32, 4096, forward_only=True)
    ref_lp = feng.forward(model, ft)
    ft2 = mtt.TokenTrie(seqs)
    ft2.forward_permute()
    o_lp = tree_oracle.tree_stack_forward(model, ft2.inputs, ft2.lcp_lens, ft2.attach_lists, len(seqs))
    assert max((a - b).abs().max().item() for a, b in zip(ref_lp, o_lp)) < 1e-5
