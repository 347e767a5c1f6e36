#!/usr/bin/env python
"""bench.py — trie tokens/s, fwd+bwd, Qwen3-0.6B on a tau2-16k-shaped merged trie (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model 0.6b|1.7b]

Our arm (default).  One "step" = one pass of the hot path over one batch of synthetic tries:
trunk forward with the tcgen05 tree-attention kernels, fused logprob/entropy head, loss,
full backward (attention bwd kernels, head bwd kernel), gradients accumulated into param.grad.
  value     = device-resident throughput: the trie layout (tokens, parent/depth/intervals, tile
              metadata) is already in HBM when the timed region starts; CUDA events, max over ranks.
  e2e       = the same metric through the public API with HOST inputs: (N > 1: partition of the global batch ->)
              TokenTrie(host tensors) -> backward_permute -> TreeTrainingEngine.backward -> float loss, i.e. host
              trie build, pinned H2D of the tokens/plan, linearisation kernels and the D2H loss read are inside
              the timed region every step.
  roofline  = tree-attention BACKWARD main kernel (the dominant kernel): algorithmic FLOPs per launch
              (pairs * Hq * D * 10) / mean launch duration from CUDA events recorded around the kernel inside
              the C library, against BOTH measured bf16 peaks of MEASURED_PEAKS.json (`frac_burst`,
              `frac_sustained`; `frac` = the sustained one: the kernel is timed inside a long power-capped step);
              `roofline.logprob` = HBM GB/s of the log-prob/entropy kernels, timed alone, against `hbm_gbs`.
  parity    = the SAME engine and kernels on the 28-layer model against the fp32 CPU oracle on the CPU leg's
              trie: loss and per-parameter |dg|/|g| (exp/compare_grads.py:42-59), bound by grad/...-bf16.txt.
  N > 1     = weak scaling: N structurally identical tries (seed-0 shape, shifted token ids), partitioned over
              ranks with the reference's LB_by_DFS_and_TM driven by a TreeTimeModel REFIT on this box
              (>= 16 timed steps -> JSONL -> add_data/fit, tree_time_model.py:14-39), one bin per rank, NCCL
              gradient all-reduce overlapped with the backward inside the timed region.  Secondaries:
              `unequal_forest` (independently seeded tries) and `strong_scaling` (a fixed 8-trie global batch
              at every N, LB_by_DFS_and_TM vs LB_by_n_tokens, exp/exp_dp.py:28-49).
Comparison legs (N = 1, after every primary number is taken; each records an error instead of raising):
  dense_baseline = BASELINE config 2's dense.py path on all sequences of the trie, same GPU;
  gpu_reference  = the reference ALGORITHM (DFS over the trie with a KV stack, HF sdpa attention, bf16;
                   oracle/tree_oracle.tree_stack_backward) on the same GPU, on a 16k-token trie of the same
                   generator (the port keeps every segment's graph, so the 67k trie does not fit), next to this
                   engine on that same trie;
  cpu_baseline   = the same port, fp32, on the host cores.
Reference arm (--impl reference).  The reference's algorithm as restated in oracle/tree_oracle.py, fp32 on the
host cores with all threads, on a bounded sample of the same generator (the reference itself is not present on
the GPU box).  The port builds one graph per root with no F1 recompute and no block-wise pops, so it is FASTER
than the reference's own engine (67.7 trie tok/s on 8 cores for the real one): conservative
for any ratio against it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

import torch  # noqa: E402

from dynamictreeattn_b200.workloads import cpu_sample_trie, synth_trie, tau2_trie  # noqa: E402

METRIC = "trie_tokens_per_sec_fwd_bwd"
UNIT = "trie tokens/s"
ATTACH = {"w_logprobs": -1.0, "w_entropy": 0.1}          # reference run.py:196
# bf16-vs-fp32 bounds of the reference's own artefact (grad/Qwen3-0.6B-TB-vs-DB-bf16.txt:6)
PARITY_TOL = {"grad_ratio_max": 1.1e-1, "grad_ratio_median": 3e-2, "loss_rel": 2e-2}


def example_loss(logprob, entropy, attach):
    """The loss of the reference's own drivers (run.py:149-152)."""
    return attach["w_logprobs"] * logprob.mean() + attach["w_entropy"] * entropy.mean()


def qwen3_config(name: str = "0.6b"):
    """Qwen3-0.6B / Qwen3-1.7B shapes (same heads, hidden 1024/2048, MLP 3072/6144)."""
    from transformers import Qwen3Config
    hidden, inter = {"0.6b": (1024, 3072), "1.7b": (2048, 6144)}[name]
    return Qwen3Config(vocab_size=151936, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=28,
                       num_attention_heads=16, num_key_value_heads=8, head_dim=128,
                       max_position_embeddings=40960, rms_norm_eps=1e-6, rope_theta=1000000.0,
                       tie_word_embeddings=True)


def qwen3_0p6b_config():
    return qwen3_config("0.6b")


def tau2_mini_trie(seed: int = 0):
    """A quarter-scale tau2 trie (~16k trie tokens): the gpu_reference leg's workload."""
    return synth_trie(seed, 1536, (4, 2, 2), (256, 768), vocab=151936, max_depth=4096)


def load_ncu_traffic(kernel: str, pairs: int):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture, with its provenance; value None
    when the capture was taken on a different workload than the one being timed."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    out = {"value": None, "source": "profiles/ncu_traffic.json", "captured": None}
    try:
        d = json.loads(p.read_text())
        out["captured"] = d.get("captured")
        out["capture_file"] = d.get("capture_file")
        if int(d.get("workload_pairs", -1)) == int(pairs):
            k = d[kernel]
            out["value"] = float(k["dram_bytes_read"]) + float(k["dram_bytes_write"])
    except Exception:
        pass
    return out


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        burst = d.get("bf16_tflops", 1600.0)
        return {"burst": burst, "sustained": d.get("bf16_tflops_sustained", burst), "hbm_gbs": d.get("hbm_gbs", 6650.0),
                "source": "MEASURED_PEAKS.json"}
    return {"burst": 1600.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in Path(self.path).read_text().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 8]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({names[i] for r in rows if len(r) >= 8 for i in range(4) if "Active" in r[4 + i]
                              and "Not" not in r[4 + i]})
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                       "samples": len(sm), "power_w_max": max(float(r[3]) for r in rows if len(r) >= 8)}
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's restatement of the reference algorithm on host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, seed: int = 0, keep_model: bool = False, model_name: str = "0.6b"):
    from transformers import Qwen3ForCausalLM
    from dynamictreeattn_b200.token_trie import TokenTrie
    from oracle import tree_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = Qwen3ForCausalLM(qwen3_config(model_name)).float()
    model.config._attn_implementation = "sdpa"
    seqs = cpu_sample_trie(seed)
    times, n_tok, loss = [], 0, 0.0
    n = warmup + steps
    for it in range(n):
        trie = TokenTrie(seqs, [dict(ATTACH) for _ in seqs])
        trie.backward_permute()
        n_tok = trie.get_stats("forward")["n_tree_tokens"]
        t0 = time.perf_counter()
        loss = tree_oracle.tree_stack_backward(model, trie.inputs, trie.lcp_lens, trie.attach_lists, tree_oracle.example_loss)
        dt = time.perf_counter() - t0
        if not (keep_model and it == n - 1):
            model.zero_grad(set_to_none=True)
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    sample = (f"oracle/tree_oracle.tree_stack_backward (port of the reference algorithm; one graph per root, no F1 "
              f"recompute / block pops: faster than the reference's own engine), Qwen3-{model_name}-shaped random-init "
              f"fp32, synth trie (trunk 512, fan (2,2), seg U{{128..384}}) = {n_tok} trie tokens / "
              f"{sum(len(s) for s in seqs)} sequence tokens per step")
    return {"value": n_tok * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "sample": sample, "n_tok": n_tok, "loss": loss, "model": model if keep_model else None, "seqs": seqs}


def full_model_parity(gpu_model, engine, ref_model, ref_loss: float, seqs):
    """The engine (bf16 kernels, all layers of `gpu_model`) against the fp32 CPU oracle's loss and gradients
    (already in `ref_model.grad`) on the same trie and the same seed-0 weights, with the per-parameter metric of
    exp/compare_grads.py:42-59, at the benchmarked model depth.  Three comparisons, as SURVEY.md section 8c states
    the tolerance:
      tree_vs_fp32        this engine (bf16) vs the fp32 oracle;
      dense_bf16_vs_fp32  the bf16 NOISE FLOOR of the same model on the same data: the per-sequence dense path
                          through stock HF attention (dense.py) in bf16 on this GPU vs the same fp32 oracle;
      tree_vs_dense_bf16  the reference's OWN recorded comparison (tree-bf16 vs dense-bf16, run.py:245-261,
                          grad/Qwen3-0.6B-TB-vs-DB-bf16.txt: max 1.06e-1, median 2.55e-2).
    ok = loss within 2e-2 relative; every parameter's tree_vs_fp32 <= max(1.1e-1, 2 x its dense_bf16_vs_fp32);
         tree_vs_dense_bf16: max <= 1.1e-1 and median <= 3e-2."""
    from dynamictreeattn_b200 import dense
    from dynamictreeattn_b200.token_trie import TokenTrie
    ref_grads = {n: p.grad for n, p in ref_model.named_parameters() if p.grad is not None}

    def ratios_vs(base: dict, grads: dict) -> dict:
        return {n: float((grads[n] - base[n]).norm() / (base[n].norm() + 1e-6)) for n in base if n in grads}

    def summary(r: dict) -> dict:
        v = sorted(r.values())
        worst = max(r, key=r.get)
        return {"max": v[-1], "median": v[len(v) // 2], "min": v[0], "worst_param": worst}

    gpu_model.zero_grad(set_to_none=True)
    trie = TokenTrie(seqs, [dict(ATTACH) for _ in seqs])
    trie.backward_permute()
    loss = engine.backward(gpu_model, trie, example_loss, block_size=2048)
    n_tok = engine.last_layout.n_nodes
    g_tree = {n: p.grad.float().cpu() for n, p in gpu_model.named_parameters() if p.grad is not None}
    gpu_model.zero_grad(set_to_none=True)
    loss_dense = dense.backward(gpu_model, list(seqs), [dict(ATTACH) for _ in seqs], example_loss, act_ckpt=False,
                                use_tqdm=False)
    g_dense = {n: p.grad.float().cpu() for n, p in gpu_model.named_parameters() if p.grad is not None}
    gpu_model.zero_grad(set_to_none=True)
    r_tree = ratios_vs(ref_grads, g_tree)
    r_floor = ratios_vs(ref_grads, g_dense)
    r_td = ratios_vs(g_dense, g_tree)
    over = {n: (r_tree[n], r_floor[n]) for n in r_tree if r_tree[n] > max(PARITY_TOL["grad_ratio_max"], 2.0 * r_floor[n])}
    out = {"loss": loss, "loss_ref": ref_loss, "loss_dense_bf16": loss_dense,
           "loss_rel": abs(loss - ref_loss) / max(abs(ref_loss), 1e-12),
           "tree_vs_fp32": summary(r_tree), "dense_bf16_vs_fp32": summary(r_floor), "tree_vs_dense_bf16": summary(r_td),
           "grad_ratio_max": summary(r_tree)["max"], "grad_ratio_median": summary(r_tree)["median"],
           "params_over_tolerance": len(over), "n_params": len(r_tree), "trie_tokens": n_tok,
           "metric": "||g_a - g_b|| / (||g_b|| + 1e-6) per parameter (exp/compare_grads.py:53-57)", "tolerance": PARITY_TOL}
    out["ok"] = bool(out["loss_rel"] <= PARITY_TOL["loss_rel"] and not over
                     and out["tree_vs_dense_bf16"]["max"] <= PARITY_TOL["grad_ratio_max"]
                     and out["tree_vs_dense_bf16"]["median"] <= PARITY_TOL["grad_ratio_median"])
    return out


def dense_gpu_leg(model, seqs, attachs, dev, n_seq=None):
    """BASELINE config 2's comparison leg: the per-sequence dense-causal path (`dynamictreeattn_b200.dense.backward`,
    the reference's dense.py:26-53 — stock HF attention, no prefix sharing) on the sequences of the same workload
    (all of them unless `n_seq`), fwd+bwd, on this GPU.  Reported in the reference's own convention, sequence
    tokens/s (run_all.py:156-159), next to the tree path's `config.seq_tokens_per_sec`.  Never raises."""
    try:
        from dynamictreeattn_b200 import dense
        ss = list(seqs) if n_seq is None else [seqs[i] for i in range(min(n_seq, len(seqs)))]
        aa = [dict(attachs[i]) for i in range(len(ss))]
        dense.backward(model, ss[:1], aa[:1], example_loss)            # warm-up (allocator, kernel selection)
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dense.backward(model, ss, aa, example_loss)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        toks = sum(int(s.numel()) for s in ss)
        return {"value": toks / (ms / 1e3), "unit": "sequence tokens/s", "ms": ms, "sequences": len(ss),
                "sequence_tokens": toks, "attn_implementation": str(getattr(model.config, "_attn_implementation", "?")),
                "sample": f"{len(ss)} of {len(seqs)} sequences of the bench trie, one causal fwd+bwd each"}
    except Exception as e:  # noqa: BLE001  (a comparison leg must not take the bench line down)
        return {"value": None, "unit": "sequence tokens/s", "error": f"{type(e).__name__}: {str(e)[:200]}"}
    finally:
        model.zero_grad(set_to_none=True)
        if torch.cuda.is_available():
            torch.cuda.empty_cache()


def gpu_reference_leg(model, engine, dev):
    """SURVEY.md section 8d Timing (ii): the reference tree ALGORITHM on the same GPU — DFS over the trie with a KV
    stack, one HF causal forward per trie segment against the cached prefix (HF `sdpa`), gradients stitched through
    the cached K/V — as restated in oracle/tree_oracle.tree_stack_backward (the reference itself cannot travel to the
    box), bf16, next to this engine on the SAME trie.  Never raises."""
    from dynamictreeattn_b200.token_trie import TokenTrie
    from oracle import tree_oracle
    prev = model.config._attn_implementation
    try:
        seqs = tau2_mini_trie(0)
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()

        def build():
            t = TokenTrie(seqs, [dict(ATTACH) for _ in seqs])
            t.backward_permute()
            return t

        def timed(fn):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize(dev)
            return time.perf_counter() - t0, out

        trie = build()
        n_tok = trie.get_stats("forward")["n_tree_tokens"]
        model.config._attn_implementation = "sdpa"
        ref = lambda: tree_oracle.tree_stack_backward(model, [t.to(dev) for t in trie.inputs], trie.lcp_lens,  # noqa: E731
                                                      trie.attach_lists, tree_oracle.example_loss)
        timed(ref)                                               # warm-up
        model.zero_grad(set_to_none=True)
        t_ref, loss_ref = timed(ref)
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        model.config._attn_implementation = prev
        ours = lambda: engine.backward(model, build(), example_loss, block_size=2048)   # noqa: E731
        timed(ours)
        model.zero_grad(set_to_none=True)
        t_ours, loss_ours = timed(ours)
        return {"value": n_tok / t_ref, "unit": UNIT, "ms": 1e3 * t_ref, "kind": "port", "attn_implementation": "sdpa",
                "dtype": "bf16", "trie_tokens": n_tok, "sequence_tokens": sum(int(s.numel()) for s in seqs),
                "ours_same_trie": {"value": n_tok / t_ours, "unit": UNIT, "ms": 1e3 * t_ours},
                "speedup_same_trie": t_ref / t_ours, "loss_ref": loss_ref, "loss_ours": loss_ours,
                "sample": "quarter-scale tau2 trie (trunk 1536, fan-out 4x2x2, segments U{256..768}); the port keeps every "
                          "segment's graph until the root's backward, so the 67k-token bench trie does not fit"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "error": f"{type(e).__name__}: {str(e)[:200]}"}
    finally:
        model.config._attn_implementation = prev
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()


def logprob_roofline(dev, hbm_gbs: float, rows: int = 8192, vocab: int = 151936, iters: int = 5):
    """HBM GB/s of the log-prob/entropy kernels on one head chunk of bf16 logits (the engine's shape: several labels
    per row through `row_of`), each timed alone with CUDA events: forward = 2*R*V bytes read, backward = 4*R*V
    (read + in-place write) — SURVEY.md section 8d."""
    from dynamictreeattn_b200.vocab_parallel import logprob_entropy_rows, logprob_entropy_rows_backward
    g = torch.Generator(device=dev).manual_seed(0)
    x = (torch.randn(rows, vocab, device=dev, generator=g) * 2).bfloat16()
    labels = torch.randint(0, vocab, (rows,), device=dev, generator=g)
    row_of = torch.arange(rows, device=dev, dtype=torch.int32)
    g_lp = torch.randn(rows, device=dev, generator=g)
    g_ent = torch.randn(rows, device=dev, generator=g)
    out = torch.empty_like(x)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    f_ms, b_ms = [], []
    for it in range(iters + 2):
        a, b, c = ev(), ev(), ev()
        a.record()
        xl, lz, ent = logprob_entropy_rows(x, labels, row_of, 1.0)
        b.record()
        logprob_entropy_rows_backward(x, labels, row_of, 1.0, 0, lz, ent, g_lp, g_ent, out=out)
        c.record()
        torch.cuda.synchronize(dev)
        if it >= 2:
            f_ms.append(a.elapsed_time(b))
            b_ms.append(b.elapsed_time(c))
    nv = rows * vocab
    fwd = 2.0 * nv / (statistics.median(f_ms) / 1e3) / 1e9
    bwd = 4.0 * nv / (statistics.median(b_ms) / 1e3) / 1e9
    both = 6.0 * nv / ((statistics.median(f_ms) + statistics.median(b_ms)) / 1e3) / 1e9
    return {"bound": "hbm", "unit": "GB/s", "peak": hbm_gbs, "fwd": fwd, "bwd": bwd, "achieved": both,
            "frac": both / hbm_gbs, "frac_fwd": fwd / hbm_gbs, "frac_bwd": bwd / hbm_gbs,
            "rows": rows, "vocab": vocab, "bytes_per_row": 6 * vocab,
            "note": "row_stats + label gather + finalize (fwd) and rowsum + bwd + scatter (bwd), launches included"}


def vocab_parallel_leg(dev, hbm_gbs: float, rows: int = 16384, vocab: int = 151936, iters: int = 10):
    """BASELINE config 4: vocab-parallel log-prob + entropy over the ranks of this run (tp = world size).  Every
    rank holds the [rows, vocab / tp] bf16 shard of the same seeded logits (~ N(0, 4)); one all-gather of 16 B per
    row and rank in the forward, none in the backward.  Returns errors against the unsharded kernels (tp = 1) and
    against the fp32 definition, and HBM GB/s per rank over fwd + bwd (6 * rows * vocab / tp bytes), timed with the
    exchange inside, max over ranks.  Collective: every rank must call it."""
    import torch.distributed as dist
    from dynamictreeattn_b200.vocab_parallel import gather_logprobs, gather_logprobs_entropy
    from oracle import tree_oracle
    world, rank = dist.get_world_size(), dist.get_rank()
    vs = vocab // world
    g = torch.Generator(device=dev).manual_seed(0)
    logits = (torch.randn(rows, vocab, device=dev, generator=g) * 2).bfloat16()
    labels = torch.randint(0, vocab, (rows,), device=dev, generator=g)
    w1 = torch.randn(rows, device=dev, generator=g)
    w2 = torch.randn(rows, device=dev, generator=g)
    shard = logits[:, rank * vs:(rank + 1) * vs].contiguous().requires_grad_(True)
    lp, ent = gather_logprobs_entropy(shard, labels, temperature=0.9, tp_group=dist.group.WORLD)
    ((lp * w1).sum() + (ent * w2).sum()).backward()
    full = logits.clone().requires_grad_(True)
    lp1, ent1 = gather_logprobs_entropy(full, labels, temperature=0.9)          # tp = 1: the unsharded kernels
    ((lp1 * w1).sum() + (ent1 * w2).sum()).backward()
    e_tp1 = (lp - lp1).abs().max().item()
    e_ent_tp1 = (ent - ent1).abs().max().item()
    g_tp1 = full.grad[:, rank * vs:(rank + 1) * vs]
    e_g_tp1 = ((shard.grad.float() - g_tp1.float()).norm() / g_tp1.float().norm()).item()
    del full, g_tp1
    e_lp = e_ent = e_g = 0.0
    for a in range(0, rows, 2048):                                              # fp32 definition, block of rows at a time
        ref = logits[a:a + 2048].float().requires_grad_(True)
        lpr, entr = tree_oracle.logprob_entropy(ref, labels[a:a + 2048], 0.9)
        ((lpr * w1[a:a + 2048]).sum() + (entr * w2[a:a + 2048]).sum()).backward()
        e_lp = max(e_lp, (lp[a:a + 2048] - lpr).abs().max().item())
        e_ent = max(e_ent, (ent[a:a + 2048] - entr).abs().max().item())
        gs = ref.grad[:, rank * vs:(rank + 1) * vs]
        e_g = max(e_g, ((shard.grad[a:a + 2048].float() - gs).norm() / gs.norm()).item())
        del ref, lpr, entr, gs
    lp_only = gather_logprobs(shard.detach(), labels, temperature=0.9, tp_group=dist.group.WORLD)
    e_lp_only = (lp_only - lp.detach()).abs().max().item()
    del logits
    torch.cuda.empty_cache()
    shard.grad = None
    run = lambda: sum(gather_logprobs_entropy(shard, labels, tp_group=dist.group.WORLD)).sum().backward()   # noqa: E731
    for _ in range(3):
        run()
        shard.grad = None
    torch.cuda.synchronize(dev)
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run()
        shard.grad = None
    b.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e_tp1, e_g_tp1, e_lp, e_ent, e_g, e_lp_only, a.elapsed_time(b) / iters, e_ent_tp1], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e_tp1, e_g_tp1, e_lp, e_ent, e_g, e_lp_only, ms, e_ent_tp1 = t.tolist()
    gbs = 6.0 * rows * vs / (ms * 1e-3) / 1e9
    return {"tp": world, "rows": rows, "vocab": vocab, "vocab_per_rank": vs, "fwd_bwd_ms": ms, "gbs_per_rank": gbs,
            "peak": hbm_gbs, "frac": gbs / hbm_gbs, "lp_err_vs_tp1": e_tp1, "entropy_err_vs_tp1": e_ent_tp1, "grad_rel_vs_tp1": e_g_tp1,
            "lp_err_vs_fp32": e_lp, "entropy_err_vs_fp32": e_ent, "grad_rel_vs_fp32": e_g, "lp_only_err": e_lp_only,
            "ok": bool(e_tp1 < 1e-5 and e_ent_tp1 < 5e-5 and e_lp < 1e-5 and e_ent < 5e-5 and e_g < 6e-3 and e_lp_only < 1e-5),
            "collectives_per_forward": 1, "exchange_bytes_per_row_per_rank": 16,
            "tolerance": "log-prob <= 1e-5, entropy <= 5e-5 on identical bf16 logits; gradient rel-L2 <= 6e-3 (one bf16 rounding)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)   # bounded: each step is ~1 min of CPU work
    r = cpu_reference_run(steps, warmup, model_name=args.model)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"qwen3-{args.model} tau2-16k-shaped merged trie, fwd+bwd (bounded CPU sample)",
                   "inputs_vs_l2": "n/a (host)"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def calibrate_time_model(engine, model, dev, seeds=(0, 1), sizes=(16, 12, 8, 6, 4, 3, 2, 1)):
    """Refit the reference's cost model the reference's way (tree_time_model.py:14-39, data_parallel.py:162-169):
    time >= 16 steps of varied tries on this GPU, write one JSONL record per step (run_all.py:162-165 schema:
    _get_stats features + loss, time, name), read the file back and add_data -> NNLS fit."""
    from dynamictreeattn_b200 import io_formats
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_time_model import TreeTimeModel
    rows = []
    cases = []
    for sd in seeds:
        seqs = tau2_trie(sd)
        for k in sizes:
            cases.append((f"tau2_seed{sd}_first{k}", seqs[:k]))
    # warm-up on the largest case (allocator, kernel attributes), untimed
    for name, seqs in cases[:1] + cases:
        trie = TokenTrie(seqs, [dict(ATTACH) for _ in seqs])
        trie.backward_permute()
        stats = trie.get_stats("backward", 2048)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()                                 # the reference's convention: run.py:7-9
        loss = engine.backward(model, trie, example_loss, block_size=2048)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        model.zero_grad(set_to_none=True)
        rows.append({**stats, "loss": loss, "time": dt, "name": name, "n_sequences": trie.n_sequences,
                     "n_tokens": trie.n_tokens})
    rows = rows[1:]
    with tempfile.NamedTemporaryFile("w", suffix=".jsonl", delete=False) as f:
        path = f.name
    io_formats.write_stats_jsonl(path, rows)
    tm = TreeTimeModel()
    tm.add_data(io_formats.read_stats_jsonl(path))
    os.unlink(path)
    mse = tm.fit()
    rel = [abs(float(tm.pred(r)) - r["time"]) / r["time"] for r in rows]
    info = {"samples": len(rows), "coeffs": [float(c) for c in tm.coeffs], "mse": mse,
            "rel_err_mean": statistics.mean(rel), "rel_err_max": max(rel),
            "features": ["n_leaf_sequences", "n_tree_tokens", "n_f1_tokens", "sum_prefix_len", "sum_depth"],
            "time_range_s": [min(r["time"] for r in rows), max(r["time"] for r in rows)]}
    return tm, info


def run_ours(args):
    import numpy as np
    import torch.distributed as dist
    from transformers import Qwen3ForCausalLM
    from dynamictreeattn_b200 import _cabi, attention
    from dynamictreeattn_b200.data_parallel import dp_backward, partition
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load(build_if_missing=False)
    peaks = load_peaks()

    cfg = qwen3_config(args.model)
    torch.manual_seed(0)
    model = Qwen3ForCausalLM(cfg).to(dev, torch.bfloat16)
    model.train()
    engine = TreeTrainingEngine(cfg, dev, torch.bfloat16, 16384)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MAX)     # noqa: E731
    sum_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.SUM)     # noqa: E731

    # ---- cost model: refit on this box from timed steps (every rank times the same steps; rank 0's fit is used) ----
    if args.no_calibrate:
        from dynamictreeattn_b200.tree_time_model import TreeTimeModel
        tm, tm_info = TreeTimeModel(), {"samples": 0, "note": "--no-calibrate: unfitted model (pred = n_tree_tokens)"}
    else:
        tm, tm_info = calibrate_time_model(engine, model, dev)
    if world > 1 and tm.coeffs is not None:
        c = torch.tensor(tm.coeffs, dtype=torch.float64, device=dev)
        dist.broadcast(c, 0)
        tm.coeffs = np.array(c.tolist())

    # ---- workload: one tau2-shaped trie per GPU (weak scaling); global batch partitioned by the cost model.
    # Weak scaling needs the SAME work per GPU at every N, so trie r is the seed-0 trie with its token ids
    # shifted by a per-trie constant (a bijection of the vocabulary: same structure, same pair count, disjoint
    # roots).  Independent seeds (sizes differ by ~10 %) are timed as a secondary number, see `unequal_forest`.
    base = tau2_trie(0)
    seqs_global = []
    for r in range(world):
        seqs_global += [(t + 7919 * r) % 151936 for t in base]
    attach_global = [dict(ATTACH) for _ in seqs_global]
    bins = partition(seqs_global, world, "LB_by_DFS_and_TM", tm, "backward", 2048) if world > 1 \
        else [list(range(len(seqs_global)))]
    seqs = [seqs_global[i] for i in bins[rank]]
    attachs = [dict(attach_global[i]) for i in bins[rank]]

    def build_trie(ss=None, aa=None):
        ss = seqs if ss is None else ss
        t = TokenTrie(ss, [dict(a) for a in (attachs if aa is None else aa)])
        t.backward_permute()
        return t

    trie = build_trie()
    layout = linearize_trie(trie, dev)
    n_tok_local = layout.n_nodes
    pairs_local = int(layout.pair_count.item())
    n_tok_global = int(sum_over_ranks(float(n_tok_local)))
    seq_tok_global = int(sum_over_ranks(float(trie.n_tokens)))
    Hq, D, L = cfg.num_attention_heads, cfg.head_dim, cfg.num_hidden_layers

    # ---- device-resident timing --------------------------------------------------------------------
    def step_resident():
        return dp_backward(engine, model, None, None, example_loss, trie=trie, layout=layout, block_size=2048)

    for _ in range(args.warmup):
        step_resident()
    model.zero_grad(set_to_none=True)
    sampler = ClockSampler(local_rank)
    sync_all()
    if rank == 0:
        sampler.start()
    attention.PROFILE = []
    launches0 = _cabi.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else {}
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = _cabi.launch_count - launches0
    prof = attention.PROFILE
    attention.PROFILE = None
    bwd_ms = [a.elapsed_time(b) for k, a, b in prof if k == "bwd"]
    fwd_ms = [a.elapsed_time(b) for k, a, b in prof if k == "fwd"]
    ms_per_step = ms_total / args.steps
    value = n_tok_global / (ms_per_step / 1e3)
    n_ranges_resident = engine.last_n_ranges
    model.zero_grad(set_to_none=True)

    # ---- end-to-end timing through the public API with host inputs -----------------------------------
    part_s = []

    def step_e2e():
        if world > 1:
            t0 = time.perf_counter()
            b = partition(seqs_global, world, "LB_by_DFS_and_TM", tm, "backward", 2048)
            part_s.append(time.perf_counter() - t0)
            return dp_backward(engine, model, seqs_global, attach_global, example_loss, bins=b, block_size=2048)
        return dp_backward(engine, model, seqs_global, attach_global, example_loss, block_size=2048)

    step_e2e()
    model.zero_grad(set_to_none=True)
    part_s.clear()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    sync_all()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    h2d = n_tok_local * 8 + (2 * layout.n_leaf - 1) * 4 + (layout.n_leaf + 1) * 8
    d2h = 4 + (8 if world > 1 else 0)
    model.zero_grad(set_to_none=True)

    # ---- roofline of the dominant kernel (attention backward) ----------------------------------------
    bwd_flops = pairs_local * Hq * D * 10.0
    fwd_flops = pairs_local * Hq * D * 4.0
    bwd_mean = statistics.mean(bwd_ms) if bwd_ms else float("nan")
    fwd_mean = statistics.mean(fwd_ms) if fwd_ms else float("nan")
    achieved = bwd_flops / (bwd_mean / 1e3) / 1e12
    attn_fwd_tflops = fwd_flops / (fwd_mean / 1e3) / 1e12
    attn_share = (sum(bwd_ms) + sum(fwd_ms)) / max(ev0.elapsed_time(ev1), 1e-9)
    lp_roof = None
    try:
        lp_roof = logprob_roofline(dev, peaks["hbm_gbs"])
    except Exception as e:  # noqa: BLE001
        lp_roof = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    torch.cuda.empty_cache()

    # ---- secondary (N > 1): forest of independently seeded tries, balanced by the reference's partitioner ----
    def time_bins(seqs_all, bins_all, steps, warm):
        mine = [seqs_all[i] for i in bins_all[rank]]
        tok = 0
        t_u = None
        if mine:
            t_u = TokenTrie(mine, [dict(ATTACH) for _ in mine])
            t_u.backward_permute()
            tok = t_u.get_stats("forward")["n_tree_tokens"]
        run = lambda: dp_backward(engine, model, None, None, example_loss, trie=t_u, block_size=2048)   # noqa: E731
        for _ in range(warm):
            run()
        model.zero_grad(set_to_none=True)
        sync_all()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(steps):
            run()
        u1.record()
        sync_all()
        ms = max_over_ranks(u0.elapsed_time(u1)) / steps
        model.zero_grad(set_to_none=True)
        return ms, int(sum_over_ranks(float(tok))), int(max_over_ranks(float(tok)))

    unequal = None
    if world > 1:
        seqs_u = []
        for sd in range(world):
            seqs_u += tau2_trie(sd)
        bins_u = partition(seqs_u, world, "LB_by_DFS_and_TM", tm, "backward", 2048)
        ms_u, tok_u, tok_max = time_bins(seqs_u, bins_u, args.steps, 2)
        unequal = {"value": tok_u / (ms_u / 1e3), "unit": UNIT, "ms_per_step": ms_u, "trie_tokens": tok_u,
                   "max_rank_trie_tokens": tok_max, "seeds": list(range(world)), "partition": "LB_by_DFS_and_TM"}

    # ---- secondary: STRONG scaling.  A fixed global batch (8 tau2 tries, seeds 0..7 = 128 sequences) at every N;
    # the reference's trie-aware contiguous partitioner against its token-count LPT (exp/exp_dp.py:28-49) ----------
    strong = None
    if not args.no_strong:
        try:
            seqs_s = []
            for sd in range(8):
                seqs_s += tau2_trie(sd)
            seq_tok_s = sum(int(s.numel()) for s in seqs_s)
            strong = {"global_batch": "8 tau2 tries (seeds 0..7), 128 sequences", "sequence_tokens": seq_tok_s,
                      "unit": "sequence tokens/s", "methods": {}}
            for method in (("LB_by_DFS_and_TM", "LB_by_n_tokens") if world > 1 else ("single",)):
                b = partition(seqs_s, world, method, tm, "backward", 2048) if world > 1 else [list(range(len(seqs_s)))]
                ms_s, tok_s, tok_max_s = time_bins(seqs_s, b, 1 if world == 1 else 2, 0 if world == 1 else 1)
                strong["methods"][method] = {"ms_per_step": ms_s, "value": seq_tok_s / (ms_s / 1e3),
                                             "trie_tokens_processed": tok_s, "max_rank_trie_tokens": tok_max_s,
                                             "trie_tokens_per_sec": tok_s / (ms_s / 1e3)}
        except Exception as e:  # noqa: BLE001
            strong = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            model.zero_grad(set_to_none=True)

    # ---- N > 1: gradient parity of the data-parallel step on the hardware (SURVEY config 3) ----------------
    dp_parity = None
    if world > 1:
        # sum over ranks of per-bin gradients (NCCL, overlapped reducer) vs rank 0 running the whole batch alone
        small = []
        for sd in range(world):
            small += synth_trie(100 + sd, 512, (2, 2), (128, 384), vocab=151936, max_depth=16384)
        b = partition(small, world, "LB_by_DFS_and_TM", tm, "backward", 2048)
        model.zero_grad(set_to_none=True)
        loss_dp = dp_backward(engine, model, small, [dict(ATTACH) for _ in small], example_loss, bins=b, block_size=2048)
        g_dp = [p.grad.float().clone() for p in model.parameters()]
        model.zero_grad(set_to_none=True)
        t_all = TokenTrie(small, [dict(ATTACH) for _ in small])
        t_all.backward_permute()
        loss_1 = engine.backward(model, t_all, example_loss, block_size=2048)       # every rank, no collective
        ratios = sorted(float((a - p.grad.float()).norm() / (p.grad.float().norm() + 1e-6))
                        for a, p in zip(g_dp, model.parameters()))
        model.zero_grad(set_to_none=True)
        dp_parity = {"loss_dp": loss_dp, "loss_single": loss_1, "loss_rel": abs(loss_dp - loss_1) / max(abs(loss_1), 1e-12),
                     "grad_ratio_max": max_over_ranks(ratios[-1]), "grad_ratio_median": ratios[len(ratios) // 2],
                     "what": "sum over ranks of per-bin gradients (NCCL all-reduce) vs the whole batch on one GPU, bf16 "
                             "both; differences are bf16 summation order only"}
        del g_dp

    # ---- N > 1: BASELINE config 4, vocab-parallel log-prob / entropy with tp = N ----------------------------
    vocab_par = None
    if world > 1 and not args.no_vocab_parallel:
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        try:
            vocab_par = vocab_parallel_leg(dev, peaks["hbm_gbs"])
        except Exception as e:  # noqa: BLE001
            vocab_par = {"ok": False, "error": f"{type(e).__name__}: {str(e)[:200]}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_mem_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    dense_line = gpu_ref = None
    if world == 1 and not args.no_dense_baseline:
        dense_line = dense_gpu_leg(model, seqs, attachs, dev)
    if world == 1 and not args.no_gpu_reference:
        gpu_ref = gpu_reference_leg(model, engine, dev)
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(1, 0, keep_model=True, model_name=args.model)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        try:
            parity = full_model_parity(model, engine, r["model"], r["loss"], r["seqs"])
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": f"{type(e).__name__}: {str(e)[:200]}"}

    pair_env = os.environ.get("DTA_BWD_PAIR", "auto")
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    pair_on = layout.bwd_whole_tiles(cfg.num_key_value_heads, n_sm) if pair_env == "auto" else pair_env not in ("", "0")
    bwd_kernel = "tree_attn_bwd_pair_kernel" if pair_on else "tree_attn_bwd_kernel"
    traffic = load_ncu_traffic(bwd_kernel, pairs_local) if world == 1 else {"value": None, "source": None, "captured": None}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"qwen3-{args.model} random-init bf16, one tau2-16k-shaped merged trie per GPU "
                        "(trunk 6144, fan-out 4x2x2, segments U{1024..3072}, depth<=16384; the same seed-0 "
                        "structure with shifted token ids on every GPU), tree fwd+bwd",
            "unequal_forest": unequal, "strong_scaling": strong, "dp_gradient_parity": dp_parity,
            "vocab_parallel": vocab_par,
            "trie_tokens": n_tok_global, "sequence_tokens": seq_tok_global, "unmasked_pairs_rank0": pairs_local,
            "layers": L, "q_heads": Hq, "kv_heads": cfg.num_key_value_heads, "head_dim": D, "hidden": cfg.hidden_size,
            "partition": "LB_by_DFS_and_TM" if world > 1 else "single",
            "time_model": tm_info, "ranges_per_step": n_ranges_resident,
            "activation_recompute": bool(getattr(engine, "last_recompute", False)),
            "inputs_vs_l2": "working set >> 126 MB L2 (no flush needed)",
            "seq_tokens_per_sec": seq_tok_global / (ms_per_step / 1e3),
        },
        "roofline": {
            "bound": "tensor", "kernel": bwd_kernel, "achieved": achieved, "peak": peaks["sustained"],
            "unit": "TFLOP/s", "frac": achieved / peaks["sustained"],
            "frac_sustained": achieved / peaks["sustained"], "frac_burst": achieved / peaks["burst"],
            "peak_burst": peaks["burst"], "peak_sustained": peaks["sustained"],
            "peak_note": "`peak`/`frac` use the SUSTAINED measured bf16 figure (the kernel is timed inside a long, "
                         "power-capped step); `frac_burst` is against the burst figure SURVEY.md section 8d names",
            "traffic": traffic["value"], "traffic_source": traffic["source"], "traffic_captured": traffic["captured"],
            "traffic_file": traffic.get("capture_file"),
            "traffic_note": "DRAM bytes per launch, ncu --set full of the same kernel on the same trie, a STATIC capture "
                            "(not measured in this run); algorithmic minimum ~1.7 GB (q,k,v,o,dO read once, dq/dk/dv + fp32 "
                            "dq accumulator written): Q/dO tiles are re-read per key tile from L2, the kernel is bound by "
                            "the SM's shared-memory data pipe (DESIGN.md 3.2), not by HBM",
            "peak_source": peaks["source"],
            "flops_per_launch": bwd_flops, "ms_per_launch": bwd_mean,
            "attn_fwd_tflops": attn_fwd_tflops, "attn_fwd_ms_per_launch": fwd_mean,
            "attn_fwd_kernel": ("tree_attn_fwd_kernel" if os.environ.get("DTA_FWD_PERSIST", "auto") in ("0",)
                                or layout.n_pad // 128 > 2048 else "tree_attn_fwd_persist_kernel"),
            "attn_fwd_frac_sustained": attn_fwd_tflops / peaks["sustained"],
            "attn_fwd_frac_burst": attn_fwd_tflops / peaks["burst"],
            "attn_fwd_bwd_tflops": (bwd_flops + fwd_flops) / ((bwd_mean + fwd_mean) / 1e3) / 1e12,
            "attn_share_of_step": attn_share,
            "logprob": lp_roof,
        },
        "parity": parity,
        "cpu_baseline": cpu,
        "dense_baseline": dense_line,
        "gpu_reference": gpu_ref,
        "e2e": {"value": n_tok_global / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "partition_ms_per_step": 1e3 * statistics.mean(part_s) if part_s else 0.0},
        "gpu_launches": launches,
        "clocks": clocks,
        "peak_mem_gb": peak_mem_gb,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--model", choices=["0.6b", "1.7b"], default="0.6b")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-calibrate", action="store_true", help="skip the 16 timed steps that refit the cost model (ncu launch lists)")
    ap.add_argument("--no-vocab-parallel", action="store_true")
    ap.add_argument("--vocab-parallel-only", action="store_true",
                    help="BASELINE config 4 alone: torchrun --nproc-per-node {2,4,8} bench.py --vocab-parallel-only")
    args = ap.parse_args()
    if args.vocab_parallel_only:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        out = vocab_parallel_leg(dev, load_peaks()["hbm_gbs"])
        if dist.get_rank() == 0:
            print(json.dumps({"metric": "vocab_parallel_logprob_entropy_gbs_per_rank", "value": out["gbs_per_rank"],
                              "unit": "GB/s", "n_gpus": dist.get_world_size(), "config": {"workload": "BASELINE config 4"},
                              "vocab_parallel": out}), flush=True)
        dist.destroy_process_group()
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
