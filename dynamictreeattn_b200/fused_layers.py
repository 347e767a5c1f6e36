"""Fused decoder-layer ops around tree attention + a Qwen3 trunk that uses them.

With attention on tcgen05, the ATen elementwise chains of the HF model code (RMSNorm with fp32
up-casts, q/k-norm, RoPE, SwiGLU, residual adds: ~40 launches per layer, HF
models/qwen3/modeling_qwen3.py:50-64, 81-83, 137-181, 263-268) cost more than attention itself.
``fused_trunk_forward`` runs the same decoder arithmetic directly on the model's own parameters with
five fused sm_100a kernels (csrc/layer_kernels.cu) and cuBLAS GEMMs; gradients land in the same
``param.grad`` tensors through ordinary autograd.  Models the fused trunk does not cover (biases,
non-default RoPE, head_dim != 128, sliding-window layers) take the generic HF-plugin path instead.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import _cabi
from .attention import tree_attention
from .linearize import TrieLayout

__all__ = ["add_rmsnorm", "qknorm_rope", "swiglu", "multi_linear", "fused_trunk_forward", "supports_fused_trunk", "rope_table"]


def _st(t: torch.Tensor) -> int:
    return _cabi.stream_ptr(t.device)


class _AddRMSNorm(torch.autograd.Function):
    """(y, s) = (RMSNorm(s) * w, s) with s = x + residual (residual optional)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor, eps: float):
        lib = _cabi.load()
        _cabi.require_cuda(x, weight)
        x = x.contiguous()
        n, h = x.shape
        y = torch.empty_like(x)
        rstd = torch.empty(n, dtype=torch.float32, device=x.device)
        if residual is not None:
            residual = residual.contiguous()
            s = torch.empty_like(x)
        else:
            s = x
        _cabi.check(lib.dta_rmsnorm_fwd(x.data_ptr(), residual.data_ptr() if residual is not None else None,
                                        weight.data_ptr(), n, h, eps, y.data_ptr(),
                                        s.data_ptr() if residual is not None else None, rstd.data_ptr(), _st(x)))
        ctx.save_for_backward(s, weight, rstd)
        ctx.has_res = residual is not None
        if residual is None:
            return y
        return y, s

    @staticmethod
    def backward(ctx, dy: torch.Tensor, ds: Optional[torch.Tensor] = None):
        lib = _cabi.load()
        s, weight, rstd = ctx.saved_tensors
        n, h = s.shape
        dy = dy.contiguous()
        if ds is not None:
            ds = ds.contiguous()
        dx = torch.empty_like(s)
        dw = torch.zeros(h, dtype=torch.float32, device=s.device)
        _cabi.check(lib.dta_rmsnorm_bwd(dy.data_ptr(), s.data_ptr(), weight.data_ptr(), rstd.data_ptr(),
                                        ds.data_ptr() if ds is not None else None, n, h, dx.data_ptr(),
                                        dw.data_ptr(), _st(s)))
        return dx, (dx if ctx.has_res else None), dw.to(weight.dtype), None


def add_rmsnorm(x, residual, weight, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(RMSNorm(x + residual) * weight, x + residual); with ``residual=None`` the sum is ``x`` itself."""
    if residual is None:
        return _AddRMSNorm.apply(x, None, weight, eps), x
    return _AddRMSNorm.apply(x, residual, weight, eps)


class _QKNormRope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, wq, wk, depth, table, eps: float):
        lib = _cabi.load()
        n, hq, d = q.shape
        hkv = k.shape[1]
        q_out = torch.empty((n, hq, d), dtype=q.dtype, device=q.device)
        k_out = torch.empty((n, hkv, d), dtype=k.dtype, device=k.device)
        rstd = torch.empty((n, hq + hkv), dtype=torch.float32, device=q.device)
        _cabi.check(lib.dta_qknorm_rope_fwd(q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0),
                                            k.stride(1), wq.data_ptr(), wk.data_ptr(), depth.data_ptr(),
                                            table.data_ptr(), n, hq, hkv, d, eps, q_out.data_ptr(),
                                            k_out.data_ptr(), rstd.data_ptr(), _st(q)))
        ctx.save_for_backward(q, k, wq, wk, depth, table, rstd)
        return q_out, k_out

    @staticmethod
    def backward(ctx, dq_out, dk_out):
        lib = _cabi.load()
        q, k, wq, wk, depth, table, rstd = ctx.saved_tensors
        n, hq, d = q.shape
        hkv = k.shape[1]
        dq_out, dk_out = dq_out.contiguous(), dk_out.contiguous()
        dq = torch.empty((n, hq, d), dtype=q.dtype, device=q.device)
        dk = torch.empty((n, hkv, d), dtype=k.dtype, device=k.device)
        dwq = torch.zeros(d, dtype=torch.float32, device=q.device)
        dwk = torch.zeros(d, dtype=torch.float32, device=q.device)
        _cabi.check(lib.dta_qknorm_rope_bwd(dq_out.data_ptr(), dk_out.data_ptr(), q.data_ptr(), q.stride(0),
                                            q.stride(1), k.data_ptr(), k.stride(0), k.stride(1), wq.data_ptr(),
                                            wk.data_ptr(), depth.data_ptr(), table.data_ptr(), rstd.data_ptr(),
                                            n, hq, hkv, d, dq.data_ptr(), dq.stride(0), dq.stride(1),
                                            dk.data_ptr(), dk.stride(0), dk.stride(1), dwq.data_ptr(),
                                            dwk.data_ptr(), _st(q)))
        return dq, dk, dwq.to(wq.dtype), dwk.to(wk.dtype), None, None, None


class _QKNormRopeAttention(torch.autograd.Function):
    """q/k-norm + RoPE followed by tree attention as ONE autograd node (they are adjacent in the model,
    HF models/qwen3/modeling_qwen3.py:263-287).  Same kernels in the forward; in the backward the attention kernel's
    fp32 dQ accumulator goes straight into the q/k-norm + RoPE backward (``dta_tree_attn_bwd_f32dq`` ->
    ``dta_qknorm_rope_bwd_f32dq``): the bf16 dq tensor, its conversion kernel and one rounding disappear.
    ``kept = (out, lse)``: a recomputed layer hands back what its first pass kept and the attention forward is skipped."""

    @staticmethod
    def forward(ctx, q, k, v, wq, wk, depth, table, eps: float, layout, scaling: float, kept):
        from .attention import tree_attention_forward
        lib = _cabi.load()
        n, hq, d = q.shape
        hkv = k.shape[1]
        q_rot = torch.empty((n, hq, d), dtype=q.dtype, device=q.device)
        k_rot = torch.empty((n, hkv, d), dtype=k.dtype, device=k.device)
        rstd = torch.empty((n, hq + hkv), dtype=torch.float32, device=q.device)
        _cabi.check(lib.dta_qknorm_rope_fwd(q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0),
                                            k.stride(1), wq.data_ptr(), wk.data_ptr(), depth.data_ptr(),
                                            table.data_ptr(), n, hq, hkv, d, eps, q_rot.data_ptr(),
                                            k_rot.data_ptr(), rstd.data_ptr(), _st(q)))
        if kept is None:
            _, _, vv, out, lse = tree_attention_forward(q_rot, k_rot, v, layout, scaling)
        else:
            from .attention import _kernel_ready
            out, lse = kept
            vv = _kernel_ready(v)
        ctx.save_for_backward(q, k, wq, wk, depth, table, rstd, q_rot, k_rot, vv, out, lse)
        ctx.layout, ctx.scaling = layout, scaling
        return out.view(out.shape) if kept is not None else out

    @staticmethod
    def backward(ctx, dout):
        from .attention import tree_attention_backward
        lib = _cabi.load()
        q, k, wq, wk, depth, table, rstd, q_rot, k_rot, v, out, lse = ctx.saved_tensors
        n, hq, d = q.shape
        hkv = k.shape[1]
        dq32, dk_rot, dv = tree_attention_backward(q_rot, k_rot, v, out, lse, dout, ctx.layout, ctx.scaling, dq_fp32=True)
        dq = torch.empty((n, hq, d), dtype=q.dtype, device=q.device)
        dk = torch.empty((n, hkv, d), dtype=k.dtype, device=k.device)
        dwq = torch.zeros(d, dtype=torch.float32, device=q.device)
        dwk = torch.zeros(d, dtype=torch.float32, device=q.device)
        _cabi.check(lib.dta_qknorm_rope_bwd_f32dq(dq32.data_ptr(), dk_rot.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1),
                                                  k.data_ptr(), k.stride(0), k.stride(1), wq.data_ptr(), wk.data_ptr(),
                                                  depth.data_ptr(), table.data_ptr(), rstd.data_ptr(), n, hq, hkv, d,
                                                  dq.data_ptr(), dq.stride(0), dq.stride(1), dk.data_ptr(), dk.stride(0),
                                                  dk.stride(1), dwq.data_ptr(), dwk.data_ptr(), _st(q)))
        return dq, dk, dv, dwq.to(wq.dtype), dwk.to(wk.dtype), None, None, None, None, None, None


def qknorm_rope_attention(q, k, v, wq, wk, depth, table, eps: float, layout, scaling: float, kept=None):
    """tree_attention(qknorm_rope(q, k), v) with the fused backward; ``kept``: see ``_QKNormRopeAttention``."""
    return _QKNormRopeAttention.apply(q, k, v, wq, wk, depth, table, eps, layout, scaling, kept)


def _qk_fusable(q, k, v) -> bool:
    """The fp32-dq backward needs 16-byte aligned views with strides that are multiples of 8 elements."""
    ok = lambda t: t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and t.data_ptr() % 16 == 0   # noqa: E731
    return FUSE_QK_ATTENTION and ok(q) and ok(k) and ok(v)


FUSE_QK_ATTENTION = os.environ.get("DTA_FUSE_QK_ATTN", "1") not in ("0", "")


def qknorm_rope(q, k, wq, wk, depth, table, eps: float):
    """q [N,Hq,128], k [N,Hkv,128] (strided views ok) -> per-head RMSNorm * w, then RoPE at ``depth``."""
    return _QKNormRope.apply(q, k, wq, wk, depth, table, eps)


class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        lib = _cabi.load()
        gate, up = gate.contiguous(), up.contiguous()
        out = torch.empty_like(gate)
        _cabi.check(lib.dta_swiglu_fwd(gate.data_ptr(), up.data_ptr(), gate.numel(), out.data_ptr(), _st(gate)))
        ctx.save_for_backward(gate, up)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _cabi.load()
        gate, up = ctx.saved_tensors
        dout = dout.contiguous()
        dg, du = torch.empty_like(gate), torch.empty_like(up)
        _cabi.check(lib.dta_swiglu_bwd(dout.data_ptr(), gate.data_ptr(), up.data_ptr(), gate.numel(), dg.data_ptr(),
                                       du.data_ptr(), _st(gate)))
        return dg, du


def swiglu(gate, up):
    return _SwiGLU.apply(gate, up)


class _MultiLinear(torch.autograd.Function):
    """(x W_0^T, x W_1^T, ...) for bias-free projections that share their input (q/k/v, gate/up).

    Same cuBLAS GEMMs as separate ``F.linear`` calls; the point is the backward: d x is accumulated inside the
    GEMM epilogues (``addmm_`` with beta = 1, fp32 accumulation, one bf16 rounding) instead of by one elementwise
    add kernel per extra consumer, which autograd inserts when the same tensor feeds several ``F.linear``s.
    """

    @staticmethod
    def forward(ctx, x, *weights):
        ctx.save_for_backward(x, *weights)
        return tuple(torch.mm(x, w.t()) for w in weights)

    @staticmethod
    def backward(ctx, *grads):
        x, *weights = ctx.saved_tensors
        dx = None
        if ctx.needs_input_grad[0]:
            for g, w in zip(grads, weights):
                if g is None:
                    continue
                if dx is None:
                    dx = torch.mm(g, w)
                else:
                    dx.addmm_(g, w)
        dws = []
        for i, g in enumerate(grads):
            dws.append(torch.mm(g.t(), x) if g is not None and ctx.needs_input_grad[1 + i] else None)
        return (dx, *dws)


def multi_linear(x, *weights):
    return _MultiLinear.apply(x, *weights)


# ---------------------------------------------------------------------------------------
# trunk
# ---------------------------------------------------------------------------------------
def rope_table(rotary_emb, n_pos: int, device) -> torch.Tensor:
    """fp32 [n_pos, 128]: cos in [:, :64], sin in [:, 64:], scaled like HF (attention_scaling)."""
    inv_freq = rotary_emb.inv_freq.to(device=device, dtype=torch.float32)
    pos = torch.arange(n_pos, device=device, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    scale = float(getattr(rotary_emb, "attention_scaling", 1.0))
    return torch.cat([freqs.cos() * scale, freqs.sin() * scale], dim=1).contiguous()


def supports_fused_trunk(model) -> bool:
    cfg = model.config
    base = getattr(model, "model", None)
    if base is None or getattr(cfg, "model_type", "") != "qwen3" or not hasattr(base, "rotary_emb"):
        return False
    hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
    if hd != 128 or cfg.hidden_size % 256 != 0 or cfg.hidden_size > 8192:
        return False
    if getattr(cfg, "attention_bias", False) or getattr(cfg, "mlp_bias", False):
        return False
    rope = getattr(cfg, "rope_parameters", None) or getattr(cfg, "rope_scaling", None) or {}
    if isinstance(rope, dict) and rope.get("rope_type", "default") not in ("default", None):
        return False
    if any(t != "full_attention" for t in getattr(cfg, "layer_types", []) or []):
        return False
    p = next(model.parameters())
    return p.is_cuda and p.dtype == torch.bfloat16


def _layer_forward(layer, x, residual, n, hq, hkv, d, depth, table, eps, attn_fn, qk_attn_fn=None):
    """One decoder layer on the model's own parameters (HF models/qwen3/modeling_qwen3.py:252-291, 81-83);
    ``attn_fn(q, k, v, scaling)`` is the attention node, ``qk_attn_fn(q_raw, k_raw, v, wq, wk, scaling)`` the combined
    q/k-norm + RoPE + attention node (used when given and the views allow it)."""
    at, mlp = layer.self_attn, layer.mlp
    h, residual = add_rmsnorm(x, residual, layer.input_layernorm.weight, eps)
    q, k, v = multi_linear(h, at.q_proj.weight, at.k_proj.weight, at.v_proj.weight)
    q, k, v = q.view(n, hq, d), k.view(n, hkv, d), v.view(n, hkv, d)
    if qk_attn_fn is not None and _qk_fusable(q, k, v):
        a = qk_attn_fn(q, k, v, at.q_norm.weight, at.k_norm.weight, at.scaling)
    else:
        q, k = qknorm_rope(q, k, at.q_norm.weight, at.k_norm.weight, depth, table, eps)
        a = attn_fn(q, k, v, at.scaling)
    o = F.linear(a.view(n, hq * d), at.o_proj.weight)
    h, residual = add_rmsnorm(o, residual, layer.post_attention_layernorm.weight, eps)
    gate, up = multi_linear(h, mlp.gate_proj.weight, mlp.up_proj.weight)
    x = F.linear(swiglu(gate, up), mlp.down_proj.weight)
    return x, residual


class _RecomputedLayer(torch.autograd.Function):
    """One decoder layer that keeps only its inputs (x, residual: 2 hidden-wide rows per token) and the attention
    output + lse (the tensors the attention BACKWARD kernel needs and the only ones that are expensive to
    reproduce), and recomputes everything else — norms, projections, RoPE, SwiGLU — when the gradient arrives.

    This is the single-pass engine's counterpart of the reference's block-wise pop (tree_training_engine.py:489-513),
    which bounds the live graph to one block: here the live graph is one LAYER of the whole trie.  The attention
    forward kernel is NOT re-run (``TreeAttnFromSaved``); the recomputation costs the layer's forward GEMMs except
    ``down_proj`` and the elementwise kernels, ~7 % of a step, against a 4-6x smaller activation footprint."""

    @staticmethod
    def forward(ctx, x, residual, layer, layout, n, hq, hkv, d, depth, table, eps):
        from .attention import tree_attention_forward
        kept = {}

        def attn_first(q, k, v, scaling):
            _, _, _, out, lse = tree_attention_forward(q, k, v, layout, scaling)
            kept["out"], kept["lse"] = out, lse
            return out

        with torch.no_grad():
            x_out, res_out = _layer_forward(layer, x, residual, n, hq, hkv, d, depth, table, eps, attn_first)
        ctx.layer, ctx.layout, ctx.dims, ctx.eps = layer, layout, (n, hq, hkv, d), eps
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, residual if residual is not None else x.new_empty(0), kept["out"], kept["lse"], depth, table)
        return x_out, res_out

    @staticmethod
    def backward(ctx, gx, gres):
        from .attention import TreeAttnFromSaved
        x, residual, out, lse, depth, table = ctx.saved_tensors
        n, hq, hkv, d = ctx.dims
        layout = ctx.layout
        xd = x.detach().requires_grad_(True)
        rd = residual.detach().requires_grad_(True) if ctx.has_res else None

        def attn_again(q, k, v, scaling):
            return TreeAttnFromSaved.apply(q, k, v, layout, scaling, out, lse)

        def qk_attn_again(q, k, v, wq, wk, scaling):
            return qknorm_rope_attention(q, k, v, wq, wk, depth, table, ctx.eps, layout, scaling, kept=(out, lse))

        with torch.enable_grad():
            x_out, res_out = _layer_forward(ctx.layer, xd, rd, n, hq, hkv, d, depth, table, ctx.eps, attn_again,
                                            qk_attn_again)
        outs, grads = [], []
        for o, g in ((x_out, gx), (res_out, gres)):
            if g is not None:
                outs.append(o)
                grads.append(g)
        torch.autograd.backward(outs, grads)            # parameter gradients accumulate into param.grad directly
        return (xd.grad, rd.grad if rd is not None else None) + (None,) * 9


def fused_trunk_forward(model, layout: TrieLayout, recompute: bool = False) -> torch.Tensor:
    """Last hidden state [N, hidden] (after the final norm) of the linearised trie.  ``recompute``: keep only each
    layer's inputs and attention output and recompute the rest in the backward (``_RecomputedLayer``)."""
    from .attention import tree_attention
    base = model.model
    cfg = model.config
    n = layout.n_nodes
    eps = float(cfg.rms_norm_eps)
    hq, hkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, 128
    depth = layout.depth[:n]
    table = rope_table(base.rotary_emb, max(layout.lens) + 1, layout.tok.device)
    x = base.embed_tokens(layout.tok)                       # [N, hidden]
    residual = None
    attn = lambda q, k, v, scaling: tree_attention(q, k, v, layout, scaling)      # noqa: E731
    qk_attn = None
    if torch.is_grad_enabled():        # the combined node only changes the BACKWARD (fp32 dQ hand-over)
        qk_attn = lambda q, k, v, wq, wk, scaling: qknorm_rope_attention(q, k, v, wq, wk, depth, table, eps, layout, scaling)   # noqa: E731
    for layer in base.layers:
        if recompute and torch.is_grad_enabled():
            x, residual = _RecomputedLayer.apply(x, residual, layer, layout, n, hq, hkv, d, depth, table, eps)
        else:
            x, residual = _layer_forward(layer, x, residual, n, hq, hkv, d, depth, table, eps, attn, qk_attn)
    h, _ = add_rmsnorm(x, residual, base.norm.weight, eps)
    return h
