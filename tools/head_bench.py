"""Fused lm_head + statistics kernel (N1, csrc/head_fwd.cu) against what it replaces: the cuBLAS bf16 GEMM that
materialises the logits, alone and followed by the row-statistics kernel.  Qwen3-0.6B / 1.7B head shapes.

    python tools/head_bench.py [--rows 67183] [--hidden 1024] [--iters 5]
"""
import argparse
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=67183)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--vocab", type=int, default=151936)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--chunk", type=int, default=8192)
    args = ap.parse_args()
    from dynamictreeattn_b200.vocab_parallel import lm_head_logprob_entropy_rows, logprob_entropy_rows
    R, K, V = args.rows, args.hidden, args.vocab
    g = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randn(R, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(V, K, device="cuda", generator=g) * (2.0 / K ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (R,), device="cuda", generator=g)
    row_of = torch.arange(R, device="cuda", dtype=torch.int32)
    ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731

    def timed(fn):
        ts = []
        for it in range(args.iters + 2):
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    def gemm_only():
        for a in range(0, R, args.chunk):
            (h[a:a + args.chunk] @ w.t())

    def gemm_stats():
        for a in range(0, R, args.chunk):
            lg = h[a:a + args.chunk] @ w.t()
            logprob_entropy_rows(lg, labels[a:a + args.chunk], row_of[:lg.shape[0]], 1.0)

    def fused():
        lm_head_logprob_entropy_rows(h, w, labels, row_of, 1.0)

    t_g, t_gs, t_f = timed(gemm_only), timed(gemm_stats), timed(fused)
    # backward: d logits of one row chunk.  (a) kept bf16 logits -> row kernel in place; (b) recompute with cuBLAS + row
    # kernel; (c) the fused recompute + backward kernel
    from dynamictreeattn_b200.vocab_parallel import lm_head_dlogits_rows, logprob_entropy_rows_backward
    C = min(args.chunk, R)
    hc, lc, rc = h[:C], labels[:C], row_of[:C]
    _, lzc, entc = lm_head_logprob_entropy_rows(hc, w, lc, rc, 1.0)
    g_lp = torch.randn(C, device="cuda", generator=g)
    g_ent = torch.randn(C, device="cuda", generator=g)
    kept = hc @ w.t()
    buf = torch.empty_like(kept)
    t_b_kept = timed(lambda: logprob_entropy_rows_backward(kept, lc, rc, 1.0, 0, lzc, entc, g_lp, g_ent, out=buf))
    t_b_recompute = timed(lambda: logprob_entropy_rows_backward(hc @ w.t(), lc, rc, 1.0, 0, lzc, entc, g_lp, g_ent, out=buf))
    t_b_fused = timed(lambda: lm_head_dlogits_rows(hc, w, lc, rc, 1.0, lzc, entc, g_lp, g_ent, buf))
    t_gemm_chunk = timed(lambda: hc @ w.t())
    flops = 2.0 * R * K * V
    # agreement of the two paths (bf16-rounded logits vs fp32 accumulators)
    xl, lz, ent = lm_head_logprob_entropy_rows(h[:4096], w, labels[:4096], row_of[:4096], 1.0)
    lg = h[:4096] @ w.t()
    xl2, lz2, ent2 = logprob_entropy_rows(lg, labels[:4096], row_of[:4096], 1.0)
    print(json.dumps({"rows": R, "hidden": K, "vocab": V, "cublas_gemm_ms": round(t_g, 3), "cublas_gemm_tflops": round(flops / t_g / 1e9, 1),
                      "gemm_plus_stats_ms": round(t_gs, 3), "fused_ms": round(t_f, 3), "fused_tflops": round(flops / t_f / 1e9, 1),
                      "bwd_chunk_rows": C, "bwd_kept_logits_ms": round(t_b_kept, 3), "bwd_cublas_recompute_ms": round(t_b_recompute, 3),
                      "bwd_fused_ms": round(t_b_fused, 3), "gemm_chunk_ms": round(t_gemm_chunk, 3),
                      "fused_vs_gemm": round(t_f / t_g, 3), "fused_vs_gemm_plus_stats": round(t_f / t_gs, 3),
                      "logz_diff_vs_bf16_logits_path": float((lz - lz2).abs().max()),
                      "entropy_diff_vs_bf16_logits_path": float((ent - ent2).abs().max())}))


if __name__ == "__main__":
    main()
