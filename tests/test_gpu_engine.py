"""-m gpu end-to-end: TokenTrie -> TreeTrainingEngine on cuda:0 (bf16 kernels) vs the fp32 CPU oracle."""
import pytest
import torch
from transformers import Qwen3Config, Qwen3ForCausalLM

from oracle import tree_oracle
from trie_cases import synth_trie

pytestmark = pytest.mark.gpu


def _models():
    cfg = Qwen3Config(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=128, max_position_embeddings=8192,
                      tie_word_embeddings=True)
    torch.manual_seed(0)
    ref = Qwen3ForCausalLM(cfg).float()
    ref.config._attn_implementation = "sdpa"
    model = Qwen3ForCausalLM(cfg)
    model.load_state_dict(ref.state_dict())
    return cfg, ref, model.to("cuda", torch.bfloat16)


def _ratios(model, ref):
    out = []
    for (n, p), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
        out.append(((p.grad.float().cpu() - pr.grad).norm() / (pr.grad.norm() + 1e-6)).item())   # compare_grads.py:53-57
    return sorted(out)


def test_engine_backward_and_forward_vs_oracle():
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    cfg, ref, model = _models()
    seqs = synth_trie(11, 96, (3, 2), (40, 120), vocab=2048)
    seqs += [seqs[0][:50].clone(), seqs[1].clone(), torch.randint(0, 2048, (70,))]     # pure prefix, duplicate, 2nd root
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    loss_ref = tree_oracle.dense_backward(ref, seqs, [dict(a) for a in attachs], tree_oracle.example_loss)
    lp_ref = tree_oracle.dense_forward(ref, seqs)
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
    trie = TokenTrie(seqs, [dict(a) for a in attachs])
    trie.backward_permute()
    loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
    r = _ratios(model, ref)
    # stated tolerance (SURVEY §8c / grad/Qwen3-0.6B-TB-vs-DB-bf16.txt): max <= 1.1e-1, median <= 3e-2
    assert abs(loss - loss_ref) < 0.02 * abs(loss_ref)
    assert r[-1] < 1.1e-1 and r[len(r) // 2] < 3e-2, (r[-1], r[len(r) // 2])
    # gradients accumulate: a second identical step doubles them
    before = {n: p.grad.clone() for n, p in model.named_parameters()}
    trie = TokenTrie(seqs, [dict(a) for a in attachs])
    trie.backward_permute()
    eng.backward(model, trie, tree_oracle.example_loss, block_size=64)
    for n, p in model.named_parameters():
        assert ((p.grad.float() - 2 * before[n].float()).norm() / (before[n].float().norm() + 1e-6)).item() < 3e-2, n
    feng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192, forward_only=True)
    ftrie = TokenTrie(seqs)
    ftrie.forward_permute()
    lps = feng.forward(model, ftrie)
    for got, want, s in zip(lps, lp_ref, seqs):
        assert got.shape == (len(s) - 1,) and got.dtype == torch.float32
        assert (got.cpu() - want).abs().max().item() < 0.15          # bf16 model vs fp32 model
        assert (got.cpu() - want).abs().mean().item() < 2e-2


def test_device_linearisation_called_and_launches_counted():
    from dynamictreeattn_b200 import _cabi
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    cfg, ref, model = _models()
    seqs = synth_trie(12, 64, (2, 2), (30, 60), vocab=2048)
    before = dict(_cabi.call_count)
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
    trie = TokenTrie(seqs, [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs])
    eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
    delta = {k: _cabi.call_count.get(k, 0) - before.get(k, 0) for k in _cabi.call_count}
    assert delta["dta_trie_linearize"] == 1 and delta["dta_tile_meta"] == 1
    assert delta["dta_tree_attn_fwd"] == 2 and delta["dta_tree_attn_bwd"] == 2        # one per layer
    assert delta["dta_logprob_row_stats"] >= 1 and delta["dta_logprob_bwd"] >= 1


def test_dense_module_matches_oracle_and_tree():
    """dense.forward/backward (reference dense.py:7-53 signatures) on the GPU vs the fp32 CPU oracle."""
    from dynamictreeattn_b200 import dense
    cfg, ref, model = _models()
    seqs = synth_trie(13, 64, (2, 2), (30, 60), vocab=2048)
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    loss_ref = tree_oracle.dense_backward(ref, seqs, [dict(a) for a in attachs], tree_oracle.example_loss)
    loss = dense.backward(model, seqs, attachs, tree_oracle.example_loss, act_ckpt=False, use_tqdm=False)
    assert abs(loss - loss_ref) < 0.02 * abs(loss_ref)
    r = _ratios(model, ref)
    assert r[-1] < 1.1e-1 and r[len(r) // 2] < 3e-2
    lps = dense.forward(model, seqs, use_tqdm=False)
    lp_ref = tree_oracle.dense_forward(ref, seqs)
    assert max((a.cpu() - b).abs().max().item() for a, b in zip(lps, lp_ref)) < 0.15


def test_fused_trunk_equals_hf_plugin_path():
    """The fused-layer trunk (N3) and the generic HF-plugin trunk run the same arithmetic: same loss,
    gradients within bf16 noise of each other and both within tolerance of the fp32 oracle."""
    from dynamictreeattn_b200.fused_layers import supports_fused_trunk
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    cfg, ref, model = _models()
    assert supports_fused_trunk(model)
    seqs = synth_trie(14, 80, (3, 2), (40, 100), vocab=2048)
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    loss_ref = tree_oracle.dense_backward(ref, seqs, [dict(a) for a in attachs], tree_oracle.example_loss)
    out = {}
    for fused in (True, False):
        model.zero_grad(set_to_none=True)
        eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
        eng.use_fused_trunk = fused
        trie = TokenTrie(seqs, [dict(a) for a in attachs])
        trie.backward_permute()
        loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
        r = _ratios(model, ref)
        assert abs(loss - loss_ref) < 0.02 * abs(loss_ref), (fused, loss, loss_ref)
        assert r[-1] < 1.1e-1 and r[len(r) // 2] < 3e-2, (fused, r[-1], r[len(r) // 2])
        out[fused] = (loss, {n: p.grad.float().clone() for n, p in model.named_parameters()})
    assert abs(out[True][0] - out[False][0]) < 0.01 * abs(loss_ref)


def test_memory_bounded_ranges_on_device_and_default_budget():
    """A trie over ``max_trie_tokens`` runs as several DFS ranges on the device: loss and gradients agree with the
    single pass to bf16 accuracy; without an explicit budget the engine derives one from free HBM (far above this
    trie) and does not split."""
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    cfg, ref, model = _models()
    seqs = synth_trie(12, 200, (3, 2, 2), (60, 160), vocab=2048)
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)

    def run():
        model.zero_grad(set_to_none=True)
        trie = TokenTrie(seqs, [dict(a) for a in attachs])
        trie.backward_permute()
        loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
        return loss, {n: p.grad.float().clone() for n, p in model.named_parameters()}

    loss1, g1 = run()
    assert eng.last_n_ranges == 1 and eng._token_budget(True) > 100_000
    eng.max_trie_tokens = 600
    loss2, g2 = run()
    assert eng.last_n_ranges >= 2
    assert abs(loss1 - loss2) < 0.01 * abs(loss1)
    worst = max(((g2[n] - g1[n]).norm() / (g1[n].norm() + 1e-6)).item() for n in g1)
    assert worst < 6e-2, worst


def test_qwen3_0p6b_28_layer_step_vs_fp32_cpu_oracle():
    """The BENCHMARKED model depth: a full 28-layer Qwen3-0.6B-shaped step (hidden 1024, Hq 16 / Hkv 8, V 151,936,
    seed-0 random init, bf16, every kernel of the product path) against the fp32 CPU oracle (the reference's
    algorithm, oracle.tree_stack_backward) on the ~2.4k-token sample trie of the bench's CPU leg: loss and the
    per-parameter |dg|/|g| of exp/compare_grads.py:42-59, bound by the reference's own bf16 artefact
    (grad/Qwen3-0.6B-TB-vs-DB-bf16.txt:6,315: max 1.06e-1, median 2.55e-2).  The same function feeds bench.py's
    `parity` object."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    r = bench.cpu_reference_run(1, 0, keep_model=True)
    cfg = bench.qwen3_config("0.6b")
    torch.manual_seed(0)
    model = Qwen3ForCausalLM(cfg).to("cuda", torch.bfloat16)          # same seed-0 weights as the CPU model
    w_cpu = dict(r["model"].named_parameters())["model.layers.3.mlp.up_proj.weight"]
    w_gpu = dict(model.named_parameters())["model.layers.3.mlp.up_proj.weight"]
    assert torch.equal(w_cpu.detach().bfloat16(), w_gpu.detach().cpu())
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 16384)
    par = bench.full_model_parity(model, eng, r["model"], r["loss"], r["seqs"])
    print("full-model parity:", {k: v for k, v in par.items() if k not in ("tolerance", "metric")})
    assert par["n_params"] == 310                                      # 28 x 11 + 2, the 310 rows of the grad/ artefact
    assert par["ok"], par
