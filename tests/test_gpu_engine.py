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
    eng.fused_head = True
    trie = TokenTrie(seqs, [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs])
    eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
    delta = {k: _cabi.call_count.get(k, 0) - before.get(k, 0) for k in _cabi.call_count}
    assert delta["dta_trie_linearize"] == 1 and delta["dta_tile_meta"] == 1
    assert delta["dta_tree_attn_fwd"] == 2                                            # one per layer
    assert delta.get("dta_tree_attn_bwd", 0) + delta.get("dta_tree_attn_bwd_f32dq", 0) == 2
    # the head runs fused (N1): one statistics launch over all rows, one fused recompute + backward launch per row chunk
    assert delta["dta_lm_head_stats"] == 1 and delta["dta_lm_head_dlogits"] >= 1 and delta["dta_lm_head_label_logits"] == 1
    assert delta.get("dta_logprob_row_stats", 0) == 0 and delta.get("dta_logprob_bwd", 0) == 0
    # default policy: a training step keeps bf16 logits while they fit (faster backward), a forward-only call is fused
    auto = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
    assert auto.fused_head is None
    before = dict(_cabi.call_count)
    trie = TokenTrie(seqs, [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs])
    auto.backward(model, trie, tree_oracle.example_loss, block_size=2048)
    delta = {k: _cabi.call_count.get(k, 0) - before.get(k, 0) for k in _cabi.call_count}
    assert delta["dta_lm_head_stats_logits"] >= 1 and delta["dta_logprob_bwd"] >= 1                # fused launch keeps the logits
    assert delta.get("dta_lm_head_stats", 0) == 0 and delta.get("dta_logprob_row_stats", 0) == 0
    before = dict(_cabi.call_count)
    auto.forward(model, TokenTrie(seqs))
    delta = {k: _cabi.call_count.get(k, 0) - before.get(k, 0) for k in _cabi.call_count}
    assert delta["dta_lm_head_stats"] == 1 and delta.get("dta_logprob_row_stats", 0) == 0


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
    assert eng.last_n_ranges == 1 and eng._token_budget(True) > 100_000 and not eng.last_recompute
    from transformers import Qwen3Config as _C
    deep = TreeTrainingEngine(_C(vocab_size=151936, hidden_size=1024, intermediate_size=3072, num_hidden_layers=28,
                                 num_attention_heads=16, num_key_value_heads=8, head_dim=128), "cuda", torch.bfloat16, 16384)
    assert deep._token_budget(True, recompute=True) > 2 * deep._token_budget(True)      # 28 layers: recomputation pays
    eng.max_trie_tokens = 600
    loss2, g2 = run()
    assert eng.last_n_ranges >= 2
    assert abs(loss1 - loss2) < 0.01 * abs(loss1)
    worst = max(((g2[n] - g1[n]).norm() / (g1[n].norm() + 1e-6)).item() for n in g1)
    assert worst < 6e-2, worst


def test_fused_head_equals_materialised_logits_head():
    """N1 in the engine: the fused lm_head path (no logits in HBM) and the cuBLAS + row-kernel path over materialised
    bf16 logits give the same loss and gradients up to the bf16 rounding of the logits; both within the tolerance of
    the fp32 oracle; forward-only log-probs agree as well."""
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    cfg, ref, model = _models()
    seqs = synth_trie(14, 80, (3, 2), (40, 100), vocab=2048)
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    loss_ref = tree_oracle.dense_backward(ref, seqs, [dict(a) for a in attachs], tree_oracle.example_loss)
    out = {}
    for fused in (True, False, None):               # no logits at all / library GEMM + row kernels / fused launch that keeps logits
        model.zero_grad(set_to_none=True)
        eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
        eng.fused_head = fused
        eng.head_chunk_rows = 128                   # several row chunks in the head's backward
        trie = TokenTrie(seqs, [dict(a) for a in attachs])
        trie.backward_permute()
        loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=None)
        r = _ratios(model, ref)
        assert abs(loss - loss_ref) < 0.02 * abs(loss_ref), (fused, loss, loss_ref)
        assert r[-1] < 1.1e-1 and r[len(r) // 2] < 3e-2, (fused, r[-1], r[len(r) // 2])
        ftrie = TokenTrie(seqs)
        ftrie.forward_permute()
        out[fused] = (loss, {n: p.grad.float().clone() for n, p in model.named_parameters()}, eng.forward(model, ftrie))
    assert abs(out[True][0] - out[False][0]) < 5e-3 * abs(loss_ref)
    worst = max(((out[True][1][n] - g).norm() / (g.norm() + 1e-6)).item() for n, g in out[False][1].items())
    assert worst < 3e-2, worst
    for a, b in zip(out[True][2], out[False][2]):
        assert (a - b).abs().max().item() < 5e-2
    # the logits-keeping fused launch reproduces the library-GEMM path: same rounded logits up to the accumulation
    # order of the two GEMMs
    assert abs(out[None][0] - out[False][0]) < 1e-3 * abs(loss_ref)
    worst = max(((out[None][1][n] - g).norm() / (g.norm() + 1e-6)).item() for n, g in out[False][1].items())
    assert worst < 1e-2, worst


def test_layer_recompute_same_gradients_less_memory():
    """Per-layer activation recomputation (fused_layers._RecomputedLayer; the engine's answer to the reference's
    block-wise pop, tree_training_engine.py:489-513): same loss, the same gradients up to bf16 accumulation order,
    the attention FORWARD kernel is not launched a second time, and the peak of live memory drops."""
    from dynamictreeattn_b200 import _cabi
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    cfg = Qwen3Config(vocab_size=2048, hidden_size=512, intermediate_size=1536, num_hidden_layers=6,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=128, max_position_embeddings=8192,
                      tie_word_embeddings=True)
    torch.manual_seed(0)
    model = Qwen3ForCausalLM(cfg).to("cuda", torch.bfloat16)
    seqs = synth_trie(15, 1500, (3, 2), (300, 900), vocab=2048)
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
    res = {}
    for rec in (False, True):
        eng.recompute = rec
        model.zero_grad(set_to_none=True)
        trie = TokenTrie(seqs, [dict(a) for a in attachs])
        trie.backward_permute()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        before = dict(_cabi.call_count)
        loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
        torch.cuda.synchronize()
        calls = {"dta_tree_attn_fwd": _cabi.call_count.get("dta_tree_attn_fwd", 0) - before.get("dta_tree_attn_fwd", 0),
                 "dta_tree_attn_bwd": sum(_cabi.call_count.get(k, 0) - before.get(k, 0)
                                          for k in ("dta_tree_attn_bwd", "dta_tree_attn_bwd_f32dq"))}
        res[rec] = (loss, {n: p.grad.float().clone() for n, p in model.named_parameters()},
                    torch.cuda.max_memory_allocated() - base, calls)
        assert eng.last_recompute == rec and eng.last_n_ranges == 1
    assert res[True][3] == res[False][3] == {"dta_tree_attn_fwd": 6, "dta_tree_attn_bwd": 6}
    assert abs(res[True][0] - res[False][0]) < 1e-3 * abs(res[False][0])
    worst = max(((res[True][1][n] - g).norm() / (g.norm() + 1e-6)).item() for n, g in res[False][1].items())
    assert worst < 2e-2, worst
    assert res[True][2] < 0.6 * res[False][2], (res[True][2], res[False][2])


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
    for fused_head in (None, True):            # default head policy (kept bf16 logits here), then the fused lm_head (N1)
        eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 16384)
        eng.fused_head = fused_head
        par = bench.full_model_parity(model, eng, r["model"], r["loss"], r["seqs"])
        print(f"full-model parity (fused_head={fused_head}):", {k: v for k, v in par.items() if k not in ("tolerance", "metric")})
        assert par["n_params"] == 310                                  # 28 x 11 + 2, the 310 rows of the grad/ artefact
        assert par["ok"], par


def test_qwen3_1p7b_tau2_trie_runs_in_one_pass_with_recompute():
    """Qwen3-1.7B (hidden 2048, MLP 6144, 28 layers; the first model of the reference's sweeps, exp/exp_tree.py) on
    the 67k-token tau2 trie: full activations (~2.2 MB per token) do not fit 180 GB, so the engine switches to
    per-layer recomputation and still runs the trie in ONE pass (no DFS range splitting, no recomputed prefixes).
    Forcing recomputation off makes it fall back to ranges; both give the same loss."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    from dynamictreeattn_b200.workloads import tau2_trie
    cfg = bench.qwen3_config("1.7b")
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = Qwen3ForCausalLM(cfg).to(torch.bfloat16)
    seqs = tau2_trie(0)
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 16384)
    out = {}
    for rec in (None, False):
        eng.recompute = rec
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        trie = TokenTrie(seqs, [dict(bench.ATTACH) for _ in seqs])
        trie.backward_permute()
        loss = eng.backward(model, trie, bench.example_loss, block_size=2048)
        torch.cuda.synchronize()
        g = model.model.layers[13].mlp.down_proj.weight.grad.float().clone()
        out[rec] = (loss, eng.last_n_ranges, eng.last_recompute, torch.cuda.max_memory_allocated() / 2 ** 30, g)
        assert torch.isfinite(g).all()
    print("1.7B tau2:", {str(k): v[:4] for k, v in out.items()})
    assert out[None][1] == 1 and out[None][2] is True                 # one pass, recomputed layers
    assert out[False][1] >= 2 and out[False][2] is False              # ranges when recomputation is off
    assert abs(out[None][0] - out[False][0]) < 2e-3 * abs(out[False][0])
    assert ((out[None][4] - out[False][4]).norm() / out[False][4].norm()).item() < 5e-2
