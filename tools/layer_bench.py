"""HBM GB/s of the fused layer kernels at the bench shape (N = 67,183 nodes, Qwen3-0.6B: hidden 1024, Hq 16 / Hkv 8,
MLP 3072), each timed alone with CUDA events:  python tools/layer_bench.py [--only qknorm]"""
import argparse, json, statistics, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from dynamictreeattn_b200.fused_layers import add_rmsnorm, qknorm_rope, swiglu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=67183)
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=7)
a = ap.parse_args()
n, hq, hkv, d, hid, inter = a.n, 16, 8, 128, 1024, 3072
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()   # noqa: E731
ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731


def timed(fn):
    ts = []
    for it in range(a.iters + 2):
        s, e = ev(), ev()
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        if it >= 2:
            ts.append(s.elapsed_time(e))
    return statistics.median(ts)


out = {}
if a.only in ("", "qknorm"):
    qkv = rnd(n, (hq + 2 * hkv) * d)
    q = qkv[:, :hq * d].view(n, hq, d)                      # strided views, as the engine hands them over
    k = qkv[:, hq * d:(hq + hkv) * d].view(n, hkv, d)
    wq, wk = (1 + 0.1 * rnd(d)), (1 + 0.1 * rnd(d))
    depth = torch.randint(0, 16384, (n,), device="cuda", dtype=torch.int32, generator=g)
    inv = 1.0 / (1e6 ** (torch.arange(0, d, 2, device="cuda").float() / d))
    fr = torch.arange(16384, device="cuda").float()[:, None] * inv[None]
    table = torch.cat([fr.cos(), fr.sin()], 1).contiguous()
    t_f = timed(lambda: qknorm_rope(q, k, wq, wk, depth, table, 1e-6))
    # backward: the kernel alone, through the C ABI (the autograd node adds view / dtype glue)
    import ctypes as C
    from dynamictreeattn_b200 import _cabi
    lib = _cabi.load()
    qo, ko = qknorm_rope(q, k, wq, wk, depth, table, 1e-6)
    rstd = torch.empty((n, hq + hkv), dtype=torch.float32, device="cuda")
    lib.dta_qknorm_rope_fwd(q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1), wq.data_ptr(),
                            wk.data_ptr(), depth.data_ptr(), table.data_ptr(), n, hq, hkv, d, 1e-6, qo.data_ptr(), ko.data_ptr(),
                            rstd.data_ptr(), _cabi.stream_ptr(q.device))
    gq, gk = torch.randn_like(qo), torch.randn_like(ko)
    dq, dk = torch.empty_like(qo), torch.empty_like(ko)
    dwq, dwk = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    t_b = timed(lambda: lib.dta_qknorm_rope_bwd(gq.data_ptr(), gk.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(),
                                                k.stride(0), k.stride(1), wq.data_ptr(), wk.data_ptr(), depth.data_ptr(),
                                                table.data_ptr(), rstd.data_ptr(), n, hq, hkv, d, dq.data_ptr(), dq.stride(0),
                                                dq.stride(1), dk.data_ptr(), dk.stride(0), dk.stride(1), dwq.data_ptr(),
                                                dwk.data_ptr(), _cabi.stream_ptr(q.device)))
    bf = n * (hq + hkv) * d * 2 * 2
    bb = n * (hq + hkv) * d * 2 * 3
    out["qknorm_rope_fwd"] = {"ms": round(t_f, 4), "gbs": round(bf / t_f / 1e6, 1)}
    out["qknorm_rope_bwd"] = {"ms": round(t_b, 4), "gbs": round(bb / t_b / 1e6, 1)}
if a.only in ("", "swiglu"):
    gt, up = rnd(n, inter).requires_grad_(True), rnd(n, inter).requires_grad_(True)
    t_f = timed(lambda: swiglu(gt.detach(), up.detach()))
    o = swiglu(gt, up)
    go = torch.randn_like(o)
    t_b = timed(lambda: torch.autograd.grad(o, [gt, up], go, retain_graph=True))
    out["swiglu_fwd"] = {"ms": round(t_f, 4), "gbs": round(n * inter * 2 * 3 / t_f / 1e6, 1)}
    out["swiglu_bwd"] = {"ms": round(t_b, 4), "gbs": round(n * inter * 2 * 5 / t_b / 1e6, 1)}
if a.only in ("", "rmsnorm"):
    x, r = rnd(n, hid).requires_grad_(True), rnd(n, hid).requires_grad_(True)
    w = (1 + 0.1 * rnd(hid)).requires_grad_(True)
    t_f = timed(lambda: add_rmsnorm(x.detach(), r.detach(), w.detach(), 1e-6))
    y, s = add_rmsnorm(x, r, w, 1e-6)
    gy, gs = torch.randn_like(y), torch.randn_like(s)
    t_b = timed(lambda: torch.autograd.grad([y, s], [x, r, w], [gy, gs], retain_graph=True))
    out["add_rmsnorm_fwd"] = {"ms": round(t_f, 4), "gbs": round(n * hid * 2 * 4 / t_f / 1e6, 1)}
    out["add_rmsnorm_bwd"] = {"ms": round(t_b, 4), "gbs": round(n * hid * 2 * 4 / t_b / 1e6, 1)}
print(json.dumps(out))
