"""Engine HOST logic (edge plan, path gathers, head chunking, autograd wiring, HF plugin) on CPU.

The CUDA ops are replaced by oracle-backed stand-ins (tests/fake_backend.py): this checks everything
around the kernels; the kernels themselves are checked by the -m gpu tests.  BASELINE config 1.
"""
import json
from pathlib import Path

import pytest
import torch
from transformers import Qwen3Config, Qwen3ForCausalLM

from dynamictreeattn_b200.token_trie import TokenTrie
from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
from oracle import tree_oracle
from trie_cases import config1_seqs, random_seqs

GOLDEN = json.loads((Path(__file__).parent / "golden" / "config1_engine.json").read_text())


def tiny_model():
    cfg = Qwen3Config(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=4096,
                      tie_word_embeddings=True)
    torch.manual_seed(0)
    m = Qwen3ForCausalLM(cfg)
    m.config._attn_implementation = "sdpa"
    return cfg, m


def attachs(n):
    return [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in range(n)]


def grads(m):
    return {n: p.grad.clone() for n, p in m.named_parameters()}


def worst_ratio(a, b):
    return max(((a[n] - b[n]).norm() / (b[n].norm() + 1e-6)).item() for n in b)


def test_config1_backward_matches_reference_golden(cpu_backend):
    cfg, model = tiny_model()
    seqs = config1_seqs()
    eng = TreeTrainingEngine(cfg, "cpu", torch.float32, 4096)
    trie = TokenTrie(seqs, attachs(len(seqs)))
    trie.backward_permute()
    loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=16)
    assert abs(loss - GOLDEN["loss_tree"]) < 1e-4 and abs(loss - GOLDEN["loss_dense"]) < 1e-4
    probe = torch.Generator().manual_seed(GOLDEN["probe_seed"])
    for n, p in model.named_parameters():
        w = torch.randn(p.grad.shape, generator=probe)
        g = GOLDEN["grads"][n]
        assert abs(p.grad.norm().item() - g["norm_tree"]) <= 2e-5 * max(g["norm_tree"], 1e-3), n
        assert abs((p.grad * w).sum().item() - g["proj_tree"]) <= 2e-4 * max(abs(g["proj_tree"]), g["norm_tree"], 1e-3), n
    assert eng.last_layout.n_nodes == GOLDEN["stats_backward_16"]["n_tree_tokens"] == 352
    assert int(eng.last_layout.pair_count) == GOLDEN["stats_backward_16"]["sum_depth"] + 352


def test_config1_forward_matches_reference_golden(cpu_backend):
    cfg, model = tiny_model()
    seqs = config1_seqs()
    eng = TreeTrainingEngine(cfg, "cpu", torch.float32, 4096, forward_only=True)
    trie = TokenTrie(seqs)
    trie.forward_permute()
    lps = eng.forward(model, trie)
    assert len(lps) == len(seqs)
    for got, want, s in zip(lps, GOLDEN["logprobs"], seqs):
        assert got.dtype == torch.float32 and got.shape == (len(s) - 1,)
        assert (got - torch.tensor(want)).abs().max().item() < 2e-5


@pytest.mark.parametrize("seed", [0, 3, 7])
def test_tree_equals_dense_on_adversarial_tries(cpu_backend, seed):
    """Forests, duplicates, full-prefix sequences: tree engine == per-sequence dense oracle (fp32)."""
    cfg, model = tiny_model()
    seqs = [s for s in random_seqs(seed, 10, vocab=6, max_len=14) if s.numel() >= 2]
    model.zero_grad()
    loss_d = tree_oracle.dense_backward(model, seqs, attachs(len(seqs)), tree_oracle.example_loss)
    gd = grads(model)
    for permute in ("backward", "forward", "none"):
        model.zero_grad()
        eng = TreeTrainingEngine(cfg, "cpu", torch.float32, 4096)
        trie = TokenTrie(seqs, attachs(len(seqs)))
        if permute == "backward":
            trie.backward_permute()
        elif permute == "forward":
            trie.forward_permute()
        loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=4)
        assert abs(loss - loss_d) < 1e-4 * max(1.0, abs(loss_d))
        assert worst_ratio(grads(model), gd) < 1e-5
    model.zero_grad()
    trie = TokenTrie(seqs, attachs(len(seqs)))
    trie.backward_permute()
    loss_s = tree_oracle.tree_stack_backward(model, trie.inputs, trie.lcp_lens, trie.attach_lists, tree_oracle.example_loss)
    assert abs(loss_s - loss_d) < 1e-4 * max(1.0, abs(loss_d))
    assert worst_ratio(grads(model), gd) < 1e-5


def test_gradients_accumulate_and_chunking_is_invisible(cpu_backend):
    cfg, model = tiny_model()
    seqs = config1_seqs()
    eng = TreeTrainingEngine(cfg, "cpu", torch.float32, 4096)
    eng.head_chunk_rows = 50                      # 352 rows -> 8 ragged chunks
    trie = TokenTrie(seqs, attachs(len(seqs)))
    loss1 = eng.backward(model, trie, tree_oracle.example_loss, block_size=None)
    g1 = grads(model)
    eng.head_chunk_rows = 8192
    trie = TokenTrie(seqs, attachs(len(seqs)))
    loss2 = eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
    g2 = grads(model)                             # never zeroed: must now hold twice the gradient
    assert abs(loss1 - loss2) < 1e-5
    for n in g1:
        assert torch.allclose(g2[n], 2 * g1[n], rtol=2e-4, atol=1e-6), n


def test_max_seq_len_and_empty_trie(cpu_backend):
    cfg, model = tiny_model()
    eng = TreeTrainingEngine(cfg, "cpu", torch.float32, max_seq_len=100)
    with pytest.raises(AssertionError, match="max_seq_len"):
        eng.backward(model, TokenTrie(config1_seqs(), attachs(9)), tree_oracle.example_loss, block_size=16)
    assert eng.backward(model, TokenTrie([], []), tree_oracle.example_loss, block_size=16) == 0.0
    assert eng.forward(model, TokenTrie([])) == []


def test_custom_loss_fn_with_token_coupling(cpu_backend):
    """loss_fn is an arbitrary callable over per-sequence vectors (tree_training_engine.py:397-398)."""
    cfg, model = tiny_model()
    seqs = config1_seqs()[:5]

    def loss_fn(lp, ent, a):
        return -(torch.softmax(lp, 0) * lp).sum() * a["k"] + ent.max() * 0.01

    att = lambda: [{"k": 0.5 + i} for i in range(len(seqs))]
    model.zero_grad()
    ld = tree_oracle.dense_backward(model, seqs, att(), loss_fn)
    gd = grads(model)
    model.zero_grad()
    trie = TokenTrie(seqs, att())
    trie.backward_permute()
    lt = TreeTrainingEngine(cfg, "cpu", torch.float32, 4096).backward(model, trie, loss_fn, block_size=32)
    assert abs(lt - ld) < 1e-4 and worst_ratio(grads(model), gd) < 1e-5


def test_memory_bounded_ranges_give_the_same_result(cpu_backend):
    """A trie over the engine's token budget runs as several sub-tries (contiguous DFS ranges from the reference's
    partitioner): same loss, same gradients, same per-sequence log-probs, original sequence ids preserved."""
    cfg, model = tiny_model()
    seqs = config1_seqs()                              # 9 sequences (one pure prefix), 352 trie tokens
    seen = []

    def loss_fn(lp, ent, a):
        seen.append(a["_sequence_batch_id"])
        return tree_oracle.example_loss(lp, ent, a) * (1 + 0.1 * a["_sequence_batch_id"])

    eng = TreeTrainingEngine(cfg, "cpu", torch.float32, 4096)
    trie = TokenTrie(seqs, attachs(len(seqs)))
    trie.backward_permute()
    loss_whole = eng.backward(model, trie, loss_fn, block_size=None)
    assert eng.last_n_ranges == 1 and sorted(seen) == list(range(9))
    g_whole = grads(model)
    model.zero_grad(set_to_none=True)
    lp_whole = eng.forward(model, TokenTrie(seqs))

    seen.clear()
    eng.max_trie_tokens = 150                          # forces >= 3 ranges; shared trunk recomputed per range
    trie = TokenTrie(seqs, attachs(len(seqs)))
    trie.backward_permute()
    loss_split = eng.backward(model, trie, loss_fn, block_size=None)
    assert eng.last_n_ranges >= 3 and sorted(seen) == list(range(9))
    assert abs(loss_split - loss_whole) < 1e-4 * max(1.0, abs(loss_whole))
    assert worst_ratio(grads(model), g_whole) < 1e-5
    lp_split = eng.forward(model, TokenTrie(seqs))
    assert eng.last_n_ranges >= 3 and len(lp_split) == 9
    for a, b in zip(lp_whole, lp_split):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-5)
    # a budget below the longest sequence cannot be met: one sequence per range is the floor, not an error
    eng.max_trie_tokens = 10
    model.zero_grad(set_to_none=True)
    trie = TokenTrie(seqs, attachs(len(seqs)))
    assert abs(eng.backward(model, trie, loss_fn, block_size=None) - loss_whole) < 1e-4 * max(1.0, abs(loss_whole))
    assert eng.last_n_ranges >= 8


def test_blockwise_attention_oracle_equals_definition():
    """oracle.masked_tree_attention_grads (block-by-block evaluation used at benchmark sizes on the GPU box) is the
    same function as the one-shot definition, including ragged last blocks and GQA."""
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.workloads import synth_trie
    from oracle import trie_oracle
    trie = TokenTrie(synth_trie(1, 100, (3, 2), (40, 90), vocab=1000))
    trie.backward_permute()
    lin = trie_oracle.linearize([t.tolist() for t in trie.inputs], trie.lcp_lens)
    N = len(lin["tok"])
    se = torch.from_numpy(lin["subtree_end"])
    g = torch.Generator().manual_seed(0)
    q = torch.randn(N, 4, 16, generator=g).requires_grad_()
    k = torch.randn(N, 2, 16, generator=g).requires_grad_()
    v = torch.randn(N, 2, 16, generator=g).requires_grad_()
    do = torch.randn(N, 4, 16, generator=g)
    o, _ = tree_oracle.masked_tree_attention(q, k, v, se, 0.25)
    o.backward(do)
    o2, dq, dk, dv = tree_oracle.masked_tree_attention_grads(q.detach(), k.detach(), v.detach(), do, se, 0.25, q_chunk=97)
    for a, b in ((o, o2), (q.grad, dq), (k.grad, dk), (v.grad, dv)):
        assert (a.detach() - b).abs().max().item() < 2e-5
