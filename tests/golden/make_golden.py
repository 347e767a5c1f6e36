"""Generate tests/golden/*.json by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/make_golden.py
The reference has no golden vectors of its own; these files pin the oracle and the
product's host code to the reference's actual outputs.
"""
import json
import os
import random
import sys
import types
from pathlib import Path

os.environ["TORCHDYNAMO_DISABLE"] = "1"
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
areal = types.ModuleType("areal"); plat = types.ModuleType("areal.platforms")
plat.is_npu_available = False; areal.platforms = plat
sys.modules["areal"] = areal; sys.modules["areal.platforms"] = plat
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
import trie as rtrie  # noqa: E402
import token_trie as rtt  # noqa: E402
import data_parallel as rdp  # noqa: E402
import tree_time_model as rtm  # noqa: E402
import tree_training_engine as rte  # noqa: E402
import dense as rdense  # noqa: E402
import vocab_parallel as rvp  # noqa: E402
from trie_cases import random_seqs, config1_seqs, synth_trie  # noqa: E402
from types import SimpleNamespace  # noqa: E402


def trie_case(seqs):
    t = rtt.TokenTrie(seqs, [{} for _ in seqs])
    out = {
        "seqs": [s.tolist() for s in seqs],
        "inputs": [x.tolist() for x in t.inputs],
        "lens": t.lens, "lcp_lens": t.lcp_lens,
        "attach": [[(a["_sequence_batch_id"], l) for a, l in al] for al in t.attach_lists],
        "n_sequences": t.n_sequences, "n_tokens": t.n_tokens,
        "stats_forward": t.get_stats("forward"),
    }
    for bs in (None, 3, 2048):
        try:
            out[f"stats_backward_{bs}"] = t.get_stats("backward", bs)
        except IndexError:          # reference defect trie.py:35
            out[f"stats_backward_{bs}"] = None
    ct = rtrie.CompressedTrie(t.lens, t.lcp_lens)
    out["order_forward"] = [list(x) for x in ct.get_order_forward()]
    out["order_backward"] = [list(x) for x in ct.get_order_backward()]
    out["order_random7"] = ct.get_order_random(7)
    n = len(t.lens)
    rng = random.Random(len(seqs))
    sub = sorted(rng.sample(range(n), max(1, n // 2)))
    out["subset"] = sub
    out["subset_lens"] = [list(x) for x in ct.get_lens(set(sub))]
    out["forkpos_16"] = rte._get_forkpos(t.lens, t.lcp_lens, 16)
    out["forkpos_none"] = rte._get_forkpos(None, t.lcp_lens, None)
    parts = {}
    for K in (2, 3):
        args = SimpleNamespace(K=K, mode="forward", block_size=None)
        parts[f"n_tokens_{K}"] = rdp.LB_by_n_tokens(seqs, K)
        parts[f"tm_{K}"] = rdp.LB_by_TM(seqs, rtm.TreeTimeModel(), args)
        parts[f"dfs_tm_{K}"] = rdp.LB_by_DFS_and_TM(seqs, rtm.TreeTimeModel(), args)
    out["partitions"] = parts
    t.backward_permute()
    out["after_backward_permute"] = {"lens": t.lens, "lcp_lens": t.lcp_lens,
                                     "inputs_head": [x[:4].tolist() for x in t.inputs]}
    return out


def main():
    cases = [trie_case(random_seqs(seed, 3 + seed % 11)) for seed in range(24)]
    cases.append(trie_case(config1_seqs()))
    cases.append(trie_case(synth_trie(3, 40, (3, 2, 2), (5, 30), vocab=50)))
    (HERE / "trie_cases.json").write_text(json.dumps(cases))

    # time model: fit on synthetic stats, record coefficients and a prediction
    rng = random.Random(0)
    data = []
    for i in range(40):
        st = {"n_leaf_sequences": rng.randrange(1, 30), "n_tree_tokens": rng.randrange(100, 90000),
              "n_f1_tokens": rng.randrange(0, 5000), "sum_prefix_len": rng.randrange(0, 10 ** 6),
              "sum_depth": rng.randrange(10 ** 4, 10 ** 9)}
        st["time"] = 1e-3 * st["n_leaf_sequences"] + 2e-5 * st["n_tree_tokens"] + 3e-9 * st["sum_depth"] + rng.random() * 1e-3
        data.append(st)
    tm = rtm.TreeTimeModel()
    tm.add_data(data)
    (HERE / "time_model.json").write_text(json.dumps(
        {"data": data, "coeffs": tm.coeffs.tolist(), "pred0": float(tm.pred(data[0])),
         "pred_unfitted": rtm.TreeTimeModel().pred(data[0])}))

    # logprob / entropy (vocab_parallel.py:13-27) on seeded logits
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(6, 50, generator=g) * 3
    labels = torch.randint(0, 50, (6,), generator=g)
    lp, ent = rvp.gather_logprobs_entropy(logits.unsqueeze(0), labels.unsqueeze(0), temperature=0.7)
    lp1 = rvp.gather_logprobs(logits.unsqueeze(0), labels.unsqueeze(0))
    (HERE / "logprob.json").write_text(json.dumps(
        {"logits": logits.tolist(), "labels": labels.tolist(), "temperature": 0.7,
         "logprobs_T": lp[0].tolist(), "entropy_T": ent[0].tolist(), "logprobs": lp1[0].tolist()}))

    # config 1: tiny Qwen3, tree vs dense, fp32 CPU (BASELINE.json configs[0])
    from transformers import Qwen3Config, Qwen3ForCausalLM
    cfg = Qwen3Config(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=16, max_position_embeddings=4096,
                      tie_word_embeddings=True)
    torch.manual_seed(0)
    model = Qwen3ForCausalLM(cfg)
    model.config._attn_implementation = "sdpa"
    seqs = config1_seqs()
    attach = lambda: [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    loss_fn = lambda lp, ent, a: a["w_logprobs"] * lp.mean() + a["w_entropy"] * ent.mean()
    eng = rte.TreeTrainingEngine(cfg, "cpu", torch.float32, 4096)
    t = rtt.TokenTrie(seqs, attach()); t.backward_permute()
    model.zero_grad()
    loss_tree = eng.backward(model, t, loss_fn, block_size=16)
    g_tree = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    loss_dense = rdense.backward(model, seqs, attach(), loss_fn, False, False)
    g_dense = {n: p.grad.clone() for n, p in model.named_parameters()}
    feng = rte.TreeTrainingEngine(cfg, "cpu", torch.float32, 4096, forward_only=True)
    ft = rtt.TokenTrie(seqs); ft.forward_permute()
    lps = feng.forward(model, ft)
    probe = torch.Generator().manual_seed(5)
    grads = {}
    for n, gt in g_tree.items():
        w = torch.randn(gt.shape, generator=probe)
        grads[n] = {"norm_tree": gt.norm().item(), "norm_dense": g_dense[n].norm().item(),
                    "proj_tree": (gt * w).sum().item(), "proj_dense": (g_dense[n] * w).sum().item()}
    (HERE / "config1_engine.json").write_text(json.dumps(
        {"loss_tree": loss_tree, "loss_dense": loss_dense, "logprobs": [x.tolist() for x in lps],
         "grads": grads, "probe_seed": 5,
         "stats_backward_16": rtt.TokenTrie(seqs).get_stats("backward", 16)}))
    print("golden written:", sorted(p.name for p in HERE.glob("*.json")))


if __name__ == "__main__":
    main()
