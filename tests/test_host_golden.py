"""Host-side integer code and the oracle against golden vectors produced by the unmodified reference."""
import json
import random
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dynamictreeattn_b200 import data_parallel as dp
from dynamictreeattn_b200.token_trie import TokenTrie, _lcp_torch
from dynamictreeattn_b200.tree_time_model import TreeTimeModel
from dynamictreeattn_b200.tree_training_engine import _get_forkpos
from dynamictreeattn_b200.trie import CompressedTrie, _get_stats, _get_subtrie, tree_pair_count
from oracle import trie_oracle

GOLDEN = Path(__file__).parent / "golden"
CASES = json.loads((GOLDEN / "trie_cases.json").read_text())


def _seqs(case):
    return [torch.tensor(s, dtype=torch.long) for s in case["seqs"]]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_token_trie_matches_reference(idx):
    c = CASES[idx]
    attachs = [{} for _ in c["seqs"]]
    t = TokenTrie(_seqs(c), attachs)
    assert [x.tolist() for x in t.inputs] == c["inputs"]
    assert t.lens == c["lens"] and t.lcp_lens == c["lcp_lens"]
    assert [[[a["_sequence_batch_id"], l] for a, l in al] for al in t.attach_lists] == [list(map(list, al)) for al in c["attach"]]
    assert (t.n_sequences, t.n_tokens) == (c["n_sequences"], c["n_tokens"])
    assert [a["_sequence_batch_id"] for a in attachs] == list(range(len(attachs)))   # caller dicts tagged in place
    assert t.get_stats("forward") == c["stats_forward"]
    for bs in (None, 3, 2048):
        want = c[f"stats_backward_{bs}"]
        got = t.get_stats("backward", bs)
        if want is not None:            # None = the reference crashes there (trie.py:35), see SURVEY.md §4
            assert got == want
    ct = CompressedTrie(t.lens, t.lcp_lens)
    assert [list(x) for x in ct.get_order_forward()] == c["order_forward"]
    assert [list(x) for x in ct.get_order_backward()] == c["order_backward"]
    assert ct.get_order_random(7) == c["order_random7"]
    assert [list(x) for x in ct.get_lens(set(c["subset"]))] == c["subset_lens"]
    assert [list(x) for x in ct.get_lens(c["subset"])] == c["subset_lens"]
    assert _get_forkpos(t.lens, t.lcp_lens, 16) == c["forkpos_16"]
    assert _get_forkpos(None, t.lcp_lens, None) == c["forkpos_none"]
    t.backward_permute()
    assert t.lens == c["after_backward_permute"]["lens"]
    assert t.lcp_lens == c["after_backward_permute"]["lcp_lens"]
    assert [x[:4].tolist() for x in t.inputs] == c["after_backward_permute"]["inputs_head"]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_partitioners_match_reference(idx):
    c = CASES[idx]
    seqs = _seqs(c)
    for K in (2, 3):
        args = SimpleNamespace(K=K, mode="forward", block_size=None)
        assert dp.LB_by_n_tokens(seqs, K) == c["partitions"][f"n_tokens_{K}"]
        assert dp.LB_by_TM(seqs, TreeTimeModel(), args) == c["partitions"][f"tm_{K}"]
        assert dp.LB_by_DFS_and_TM(seqs, TreeTimeModel(), args) == c["partitions"][f"dfs_tm_{K}"]
        assert dp.partition(seqs, K, "LB_by_DFS_and_TM", mode="forward") == c["partitions"][f"dfs_tm_{K}"]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_matches_reference(idx):
    """The oracle restatement is pinned to the same golden vectors."""
    c = CASES[idx]
    leaves, members, lcps = trie_oracle.sort_and_leafize(c["seqs"])
    assert leaves == c["inputs"] and lcps == c["lcp_lens"]
    assert [[list(m) for m in ms] for ms in members] == [list(map(list, al)) for al in c["attach"]]
    lens = [len(x) for x in leaves]
    assert [list(x) for x in trie_oracle.leaf_order(lens, lcps, "forward")] == c["order_forward"]
    assert [list(x) for x in trie_oracle.leaf_order(lens, lcps, "backward")] == c["order_backward"]
    fwd = trie_oracle.stats(lens, lcps, "forward")
    assert {k: c["stats_forward"][k] for k in fwd} == fwd
    for bs in (None, 3, 2048):
        want = c[f"stats_backward_{bs}"]
        if want is not None:
            got = trie_oracle.stats(lens, lcps, "backward", bs)
            assert {k: want[k] for k in got} == got


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_linearisation_identities(idx):
    """Interval mask == parent-pointer closure; #visible pairs == sum_depth + n_tree_tokens (trie.py:8-13)."""
    c = CASES[idx]
    t = TokenTrie(_seqs(c))
    t.backward_permute()
    lin = trie_oracle.linearize([x.tolist() for x in t.inputs], t.lcp_lens)
    st = _get_stats(t.lens, t.lcp_lens, "forward")
    N = len(lin["tok"])
    assert N == st["n_tree_tokens"]
    assert (trie_oracle.ancestor_mask(lin["parent"]) == trie_oracle.interval_mask(lin["subtree_end"])).all()
    assert int(trie_oracle.interval_mask(lin["subtree_end"]).sum()) == st["sum_depth"] + N == tree_pair_count(t.lens, t.lcp_lens)
    for i, ids in enumerate(t.inputs):                      # every path spells its leaf
        assert lin["tok"][lin["paths"][i]].tolist() == ids.tolist()
        assert lin["depth"][lin["paths"][i]].tolist() == list(range(len(ids)))


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_range_subtrie_equals_subset_subtrie(idx):
    """The partitioner's shortcut for contiguous leaf ranges builds the same sub-trie as the reference's
    _get_subtrie(trie, set(range(a, b))) (data_parallel.py:67-70, trie.py:240-242)."""
    from dynamictreeattn_b200.trie import _get_range_subtrie
    c = CASES[idx]
    t = TokenTrie(_seqs(c))
    full = CompressedTrie(t.lens, t.lcp_lens)
    n = len(t.lens)
    for a in range(n):
        for b in range(a + 1, n + 1):
            x, y = _get_subtrie(full, set(range(a, b))), _get_range_subtrie(full, a, b)
            assert (x.depth, x.seq_id, x.children) == (y.depth, y.seq_id, y.children)
            assert x.get_order_backward() == y.get_order_backward()


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_native_range_stats_equal_python(idx):
    """dta_host_subtrie_stats (C++ host helper of the partitioner's probes) returns the integers of the Python code —
    sub-trie -> get_order_forward / get_order_backward -> _get_stats — for every contiguous leaf range, every mode and
    block size (incl. the cases where the reference raises IndexError, where both use the intended value 0)."""
    c = CASES[idx]
    t = TokenTrie(_seqs(c))
    full = CompressedTrie(t.lens, t.lcp_lens)
    n = len(t.lens)
    for a in range(n):
        for b in range(a + 1, n + 1):
            sub = _get_subtrie(full, set(range(a, b)))
            for mode, bs in (("forward", None), ("backward", None), ("backward", 3), ("backward", 2048)):
                _, lens, lcps = sub.get_order_forward() if mode == "forward" else sub.get_order_backward()
                want = _get_stats(lens, lcps, mode, bs)
                got = dp._native_range_stats(full, a, b, mode, bs)
                assert got is not None and got == want, (a, b, mode, bs)


def test_time_model_matches_reference():
    g = json.loads((GOLDEN / "time_model.json").read_text())
    tm = TreeTimeModel()
    assert tm.pred(g["data"][0]) == g["pred_unfitted"]
    tm.add_data(g["data"][:10])
    assert tm.coeffs is None                                # below MIN_N_DATA_POINTS
    tm.add_data(g["data"][10:])
    assert np.allclose(tm.coeffs, g["coeffs"], rtol=1e-9, atol=0)
    assert abs(float(tm.pred(g["data"][0])) - g["pred0"]) <= 1e-12 * abs(g["pred0"])
    big = TreeTimeModel()
    big.add_data(g["data"] * 40)
    assert len(big.data) == TreeTimeModel.MAX_N_DATA_POINTS


def test_error_contracts():
    a, b = torch.tensor([3, 1]), torch.tensor([1, 2])
    with pytest.raises(ValueError, match="not sorted"):
        TokenTrie([a, b], sorted=True)
    with pytest.raises(AssertionError):
        TokenTrie([a, b], [{}])
    with pytest.raises(ValueError):
        CompressedTrie([3, 2], [1, 1])
    with pytest.raises(ValueError):
        _get_stats([3], [], "sideways")
    assert _lcp_torch(torch.tensor([1, 2, 3]), torch.tensor([1, 2])) == 2
    assert _lcp_torch(torch.tensor([], dtype=torch.long), torch.tensor([1])) == 0
    # empty input and single sequence
    assert TokenTrie([]).inputs == []
    one = TokenTrie([torch.tensor([5, 6, 7])])
    assert one.lens == [3] and one.lcp_lens == []
    # backward stats on the case the reference crashes on (last leaf fits one block)
    assert _get_stats([5, 4], [2], "backward", 16)["n_f1_tokens"] == 2


def test_deep_fork_chain_needs_no_recursion():
    """2000 nested forks: the reference recurses once per fork level and hits the recursion limit."""
    depth = 2000
    lens = [depth + 1 - i for i in range(depth)]
    lcps = [depth - 1 - i for i in range(depth - 1)]
    order, olens, olcps = CompressedTrie(lens, lcps).get_order_forward()
    assert sorted(order) == list(range(depth)) and len(olcps) == depth - 1
    sub = _get_subtrie(CompressedTrie(lens, lcps), set(range(0, depth, 2)))
    assert len(sub.get_order_backward()[0]) == depth // 2


@pytest.mark.parametrize("seed", range(6))
def test_layout_pair_count_and_backward_schedule_rule(seed):
    """Host-side numbers the attention backward's schedule choice rests on: TrieLayout.n_pairs (closed form over
    lens / lcp) equals the oracle's count of visible (q, k) pairs, and the whole-tile schedule is only chosen when no
    key tile's query range can exceed half of an SM's balanced share (include/dta_b200.h, bwd_whole_tiles)."""
    import fake_backend
    from trie_cases import random_seqs, synth_trie
    seqs = random_seqs(seed, 30, vocab=4, max_len=60) if seed % 2 else synth_trie(seed, 300, (3, 2), (50, 400), vocab=500)
    trie = TokenTrie(seqs)
    trie.backward_permute()
    lay = fake_backend.linearize_trie(trie, "cpu")
    lin = trie_oracle.linearize([t.tolist() for t in trie.inputs], trie.lcp_lens)
    assert lay.n_pairs == int((lin["depth"] + 1).sum()) == int(trie_oracle.interval_mask(lin["subtree_end"]).sum())
    assert lay.n_pairs == int(lay.pair_count.item())
    for n_kv, n_sm in ((8, 148), (1, 148), (2, 4)):
        whole = lay.bwd_whole_tiles(n_kv, n_sm)
        share = (lay.n_pairs // (128 * 128)) * n_kv / n_sm
        assert whole == (lay.n_tiles <= 0.5 * share)
        if whole:                                         # then every tile's live range is at most half a share
            k_qlast = lay.k_qlast.tolist()
            assert max(k_qlast[t] - t + 1 for t in range(lay.n_tiles)) <= 0.5 * share


@pytest.mark.parametrize("seed", range(4))
def test_cta_pair_schedule_premises(seed):
    """What the CTA-pair backward kernel (csrc/attn_bwd_pair.cu) takes for granted, checked against the oracle's mask:
    a cluster walks query tiles [2j, max(k_qlast[2j], k_qlast[2j+1])]; inside that range every query tile outside a
    key tile's OWN live range [t, k_qlast[t]] holds no visible pair for that tile (so the fully masked steps add
    nothing), every tile belongs to exactly one pair, and the simulator's pair order (first appearance in k_order)
    lists every pair once."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import heapq  # noqa: F401  (bwd_schedule_sim imports it)
    from bwd_schedule_sim import pair_jobs_of
    from trie_cases import random_seqs, synth_trie
    seqs = random_seqs(seed, 40, vocab=4, max_len=90) if seed % 2 else synth_trie(seed, 200, (3, 2), (60, 300), vocab=500)
    trie = TokenTrie(seqs)
    trie.backward_permute()
    lin = trie_oracle.linearize([t.tolist() for t in trie.inputs], trie.lcp_lens)
    mask = trie_oracle.interval_mask(lin["subtree_end"])            # [q, k]
    tm = trie_oracle.tile_meta(lin["subtree_end"])
    n = mask.shape[0]
    n_tiles = len(tm["k_qlast"])
    qlast = [int(x) for x in tm["k_qlast"]]
    for j in range((n_tiles + 1) // 2):
        last = max(qlast[2 * j], qlast[2 * j + 1] if 2 * j + 1 < n_tiles else 0)
        for t in (2 * j, 2 * j + 1):
            if t >= n_tiles:
                continue
            cols = slice(t * 128, min((t + 1) * 128, n))
            for qt in range(n_tiles):
                blk = mask[qt * 128:min((qt + 1) * 128, n), cols]
                if 2 * j <= qt <= last:
                    if not (t <= qt <= qlast[t]):
                        assert not blk.any()                        # a masked step of the union range
                else:
                    assert not blk.any()                            # nothing visible outside the union range
    jobs, single_steps = pair_jobs_of(seqs, n_kv=2, group=2, overhead=0)
    assert len(jobs) == 2 * ((n_tiles + 1) // 2)
    assert 2 * sum(jobs) >= single_steps == sum(qlast[t] - t + 1 for t in range(n_tiles)) * 2 * 2
