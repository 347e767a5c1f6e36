"""GPU self-test: run every kernel family against the oracle and PRINT the error numbers.

Unlike pytest -x this never stops at the first failure and isolates each section in its own
process (a CUDA fault is sticky).  Usage on the GPU box:
    python tools/gpu_selftest.py            # all sections
    python tools/gpu_selftest.py attn_fwd   # one section, in-process
Output lines start with PASS / FAIL / INFO so they can be grepped from gpurun_out/selftest.log.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SECTIONS = ["trie", "logprob", "attn_fwd", "attn_bwd", "engine"]


def report(ok: bool, name: str, **kv):
    print(("PASS" if ok else "FAIL"), name, " ".join(f"{k}={v}" for k, v in kv.items()), flush=True)


def _tries():
    """(name, sequences) pairs covering chain, fan, deep forks, forest, ragged sizes."""
    import torch
    from trie_cases import random_seqs, synth_trie
    g = torch.Generator().manual_seed(0)
    cases = []
    cases.append(("chain_300", [torch.randint(0, 1000, (300,), generator=g)]))
    cases.append(("chain_128", [torch.randint(0, 1000, (128,), generator=g)]))
    cases.append(("fan_small", synth_trie(1, 100, (3, 2), (40, 90), vocab=1000)))
    cases.append(("deep_forks", synth_trie(2, 300, (2, 2, 2, 2), (100, 260), vocab=1000)))
    forest = synth_trie(3, 200, (2, 2), (100, 200), vocab=1000) + synth_trie(4, 150, (3,), (50, 300), vocab=1000)
    cases.append(("forest", forest))
    cases.append(("adversarial", random_seqs(5, 40, vocab=4, max_len=40)))
    cases.append(("tau2_mini", synth_trie(6, 1536, (4, 2, 2), (256, 768), vocab=5000, max_depth=4096)))
    return cases


def section_trie():
    import numpy as np
    import torch
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.linearize import linearize_trie, lcp_adjacent_device
    from oracle import trie_oracle
    for name, seqs in _tries():
        trie = TokenTrie(seqs)
        trie.backward_permute()
        lay = linearize_trie(trie, "cuda")
        torch.cuda.synchronize()
        ref = trie_oracle.linearize([t.tolist() for t in trie.inputs], trie.lcp_lens)
        N = lay.n_nodes
        ok = N == len(ref["tok"])
        for key in ("depth", "parent", "subtree_end", "leaf_of"):
            got = getattr(lay, key)[:N].cpu().numpy().astype(np.int64)
            ok &= bool((got == ref[key]).all())
        ok &= bool((lay.tok.cpu().numpy() == ref["tok"]).all())
        ok &= bool((lay.leaf_off.cpu().numpy().astype(np.int64) == ref["leaf_off"]).all())
        ok &= bool((lay.path.cpu().numpy().astype(np.int64) == np.concatenate(ref["paths"])).all())
        meta = trie_oracle.tile_meta(ref["subtree_end"])
        for key in ("kmax_end", "kmin_end", "q_nlive", "k_qlast", "q_order", "k_order"):
            ok &= bool((getattr(lay, key).cpu().numpy().astype(np.int64) == meta[key]).all())
        ok &= int(lay.pair_count.item()) == int((ref["depth"] + 1).sum())
        pad_ok = bool((lay.subtree_end[N:] == 0).all()) and bool((lay.parent[N:] == -1).all())
        report(ok and pad_ok, f"trie.linearize[{name}]", N=N, leaves=lay.n_leaf)
        # adjacent LCP kernel on the sorted, un-leafized sequences
        srt = sorted(seqs, key=lambda t: t.tolist())
        lcps, bad = lcp_adjacent_device([t.cuda() for t in srt])
        want = [trie_oracle.lcp(a.tolist(), b.tolist()) for a, b in zip(srt, srt[1:])]
        report(lcps == want and not bad, f"trie.lcp_adjacent[{name}]", n=len(srt))
    if len(seqs) > 1:
        _, bad = lcp_adjacent_device([t.cuda() for t in srt[::-1]])
        report(bool(bad) or len(set(map(lambda t: tuple(t.tolist()), srt))) == 1, "trie.lcp_adjacent.unsorted_flag")


def section_logprob():
    import torch
    from dynamictreeattn_b200.vocab_parallel import gather_logprobs, gather_logprobs_entropy
    from dynamictreeattn_b200.vocab_parallel import logprob_entropy_rows, logprob_entropy_rows_backward
    from oracle import tree_oracle
    torch.manual_seed(0)
    for dtype, V, R, tol in ((torch.float32, 1000, 37, 2e-5), (torch.float32, 151936, 9, 2e-5),
                             (torch.bfloat16, 151936, 33, 2e-5), (torch.bfloat16, 18992, 64, 2e-5),
                             (torch.bfloat16, 1003, 5, 2e-5)):
        for temp in (1.0, 0.7):
            logits = (torch.randn(R, V, device="cuda") * 2).to(dtype)
            labels = torch.randint(0, V, (R,), device="cuda")
            x = logits.clone().requires_grad_(True)
            lp, ent = gather_logprobs_entropy(x, labels, temperature=temp)
            xr = logits.float().clone().requires_grad_(True)          # same (possibly bf16-rounded) values
            lpr, entr = tree_oracle.logprob_entropy(xr, labels, temp)
            e1, e2 = (lp - lpr).abs().max().item(), (ent - entr).abs().max().item()
            w1, w2 = torch.randn(R, device="cuda"), torch.randn(R, device="cuda")
            ((lp * w1).sum() + (ent * w2).sum()).backward()
            ((lpr * w1).sum() + (entr * w2).sum()).backward()
            gref = xr.grad
            gerr = ((x.grad.float() - gref).norm() / gref.norm()).item()
            gtol = 1e-5 if dtype == torch.float32 else 6e-3
            report(e1 < tol and e2 < 5 * tol and gerr < gtol, f"logprob[{str(dtype)[6:]},V={V},R={R},T={temp}]",
                   lp_err=f"{e1:.2e}", ent_err=f"{e2:.2e}", grad_rel=f"{gerr:.2e}")
    # 3-D input, logprobs only
    logits = torch.randn(2, 5, 777, device="cuda")
    labels = torch.randint(0, 777, (2, 5), device="cuda")
    lp = gather_logprobs(logits, labels)
    ref, _ = tree_oracle.logprob_entropy(logits, labels)
    report(lp.shape == labels.shape and (lp - ref).abs().max().item() < 2e-5, "logprob.3d")
    # several labels per row (trie edges)
    R, V, M = 50, 4096, 120
    x = torch.randn(R, V, device="cuda").bfloat16()
    row_of = torch.randint(0, R, (M,), device="cuda", dtype=torch.int32)
    row_of[0] = -1
    labels = torch.randperm(V, device="cuda")[:M]
    xl, lz, ent = logprob_entropy_rows(x, labels, row_of, 1.0)
    lsm = torch.log_softmax(x.float(), -1)
    want = torch.where(row_of >= 0, x.float()[row_of.clamp(min=0).long(), labels], torch.zeros(M, device="cuda"))
    e = (xl - want).abs().max().item()
    elz = (lz - torch.logsumexp(x.float(), -1)).abs().max().item()
    report(e == 0.0 and elz < 2e-5, "logprob.rows_multi_label", x_err=e, lz_err=f"{elz:.2e}")
    g_lp = torch.randn(M, device="cuda")
    g_lp[0] = 0
    g_ent = torch.randn(R, device="cuda")
    got = logprob_entropy_rows_backward(x, labels, row_of, 1.0, 0, lz, ent, g_lp, g_ent).float()
    xr = x.float().clone().requires_grad_(True)
    lsm = torch.log_softmax(xr, -1)
    entr = -(lsm.exp() * lsm).sum(-1)
    valid = row_of >= 0
    lpr = lsm[row_of[valid].long(), labels[valid]]
    ((lpr * g_lp[valid]).sum() + (entr * g_ent).sum()).backward()
    gerr = ((got - xr.grad).norm() / xr.grad.norm()).item()
    report(gerr < 6e-3, "logprob.rows_multi_label.bwd", grad_rel=f"{gerr:.2e}")


def _attn_inputs(N, Hq, Hkv, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(N, Hkv, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(N, Hkv, 128, device="cuda", generator=g).bfloat16()
    do = torch.randn(N, Hq, 128, device="cuda", generator=g).bfloat16()
    return q, k, v, do


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def section_attn_fwd(with_bwd: bool = False):
    import torch
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.linearize import linearize_trie
    from dynamictreeattn_b200.attention import tree_attention
    from oracle import tree_oracle
    for name, seqs in _tries():
        trie = TokenTrie(seqs)
        trie.backward_permute()
        lay = linearize_trie(trie, "cuda")
        for Hq, Hkv in ((2, 1), (4, 2)) if not with_bwd else ((4, 2),):
            q, k, v, do = _attn_inputs(lay.n_nodes, Hq, Hkv, 7)
            scale = 128 ** -0.5
            if with_bwd:
                q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
            t0 = time.time()
            out = tree_attention(q, k, v, lay, scale)
            torch.cuda.synchronize()
            qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
            ref, lse = tree_oracle.masked_tree_attention(qr, kr, vr, lay.subtree_end[:lay.n_nodes].long(), scale)
            err = _rel(out, ref)
            bad = (~torch.isfinite(out.float())).sum().item()
            if not with_bwd:
                report(err < 5e-3 and bad == 0, f"attn.fwd[{name},Hq={Hq},Hkv={Hkv}]", N=lay.n_nodes,
                       rel=f"{err:.2e}", nonfinite=bad, pairs=int(lay.pair_count.item()))
                continue
            out.backward(do)
            torch.cuda.synchronize()
            ref.backward(do.float())
            eq, ek, ev = _rel(q.grad, qr.grad), _rel(k.grad, kr.grad), _rel(v.grad, vr.grad)
            report(max(eq, ek, ev) < 1e-2 and err < 5e-3, f"attn.bwd[{name},Hq={Hq},Hkv={Hkv}]", N=lay.n_nodes,
                   fwd=f"{err:.2e}", dq=f"{eq:.2e}", dk=f"{ek:.2e}", dv=f"{ev:.2e}")


def section_attn_bwd():
    section_attn_fwd(with_bwd=True)


def section_engine():
    import torch
    from transformers import Qwen3Config, Qwen3ForCausalLM
    from dynamictreeattn_b200.token_trie import TokenTrie
    from dynamictreeattn_b200.tree_training_engine import TreeTrainingEngine
    from oracle import tree_oracle
    from trie_cases import synth_trie
    cfg = Qwen3Config(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=128, max_position_embeddings=8192,
                      tie_word_embeddings=True)
    torch.manual_seed(0)
    ref_model = Qwen3ForCausalLM(cfg).float()
    ref_model.config._attn_implementation = "sdpa"
    model = Qwen3ForCausalLM(cfg)
    model.load_state_dict(ref_model.state_dict())
    model = model.to("cuda", torch.bfloat16)
    model.config._attn_implementation = "sdpa"
    seqs = synth_trie(11, 96, (3, 2), (40, 120), vocab=2048)
    seqs.append(seqs[0][:50].clone())
    attachs = [{"w_logprobs": -1.0, "w_entropy": 0.1} for _ in seqs]
    # fp32 CPU oracle (dense, per sequence)
    loss_ref = tree_oracle.dense_backward(ref_model, seqs, [dict(a) for a in attachs], tree_oracle.example_loss)
    lp_ref = tree_oracle.dense_forward(ref_model, seqs)
    eng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192)
    trie = TokenTrie(seqs, [dict(a) for a in attachs])
    trie.backward_permute()
    loss = eng.backward(model, trie, tree_oracle.example_loss, block_size=2048)
    ratios = []
    for (n, p), (_, pr) in zip(model.named_parameters(), ref_model.named_parameters()):
        ratios.append(((p.grad.float().cpu() - pr.grad).norm() / (pr.grad.norm() + 1e-6)).item())
    ratios.sort()
    report(abs(loss - loss_ref) < 0.05 * abs(loss_ref) and ratios[-1] < 1.1e-1 and ratios[len(ratios) // 2] < 3e-2,
           "engine.backward", loss=f"{loss:.4f}", loss_ref=f"{loss_ref:.4f}", grad_rel_max=f"{ratios[-1]:.2e}",
           grad_rel_median=f"{ratios[len(ratios) // 2]:.2e}")
    feng = TreeTrainingEngine(cfg, "cuda", torch.bfloat16, 8192, forward_only=True)
    ftrie = TokenTrie(seqs)
    ftrie.forward_permute()
    lps = feng.forward(model, ftrie)
    err = max((a.cpu() - b).abs().max().item() for a, b in zip(lps, lp_ref))
    report(err < 0.15, "engine.forward", max_abs_lp_err=f"{err:.3e}")


def main():
    if len(sys.argv) > 1:
        globals()["section_" + sys.argv[1]]()
        return
    rc = 0
    for s in SECTIONS:
        print(f"INFO ===== section {s}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, s], timeout=420)
            if r.returncode != 0:
                print(f"FAIL section {s} exited with {r.returncode}", flush=True)
                rc = 1
        except subprocess.TimeoutExpired:
            print(f"FAIL section {s} timed out", flush=True)
            rc = 1
        print(f"INFO section {s} took {time.time() - t0:.1f}s", flush=True)
    sys.exit(rc)


if __name__ == "__main__":
    main()
