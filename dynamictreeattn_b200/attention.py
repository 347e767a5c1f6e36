"""Tree attention op: autograd wrapper over dta_tree_attn_fwd/bwd + HF plugin point.

The model-side boundary is HuggingFace's ``AttentionInterface`` (the function slot the
reference fills with flash_attention_3 / sdpa, run.py:170-171,
HF models/qwen3/modeling_qwen3.py:273-287).  ``register()`` installs the B200 kernels
under the name ``"dta_tree"``; the engine switches ``config._attn_implementation`` to it
for the duration of a step.  HF builds no mask for unknown attention names
(HF masking_utils.py:836), so non-monotonic position ids are safe.
"""
from __future__ import annotations

import ctypes as C
import threading
from contextlib import contextmanager
from typing import Optional

import torch

from . import _cabi
from .linearize import TrieLayout

ATTN_NAME = "dta_tree"
_tls = threading.local()

# bench.py sets this to a list to collect ("fwd"|"bwd", start_event, stop_event) around the MAIN
# attention kernel of every call (events are recorded inside the C library, on the launch stream).
PROFILE = None


def _arm_profile(lib, kind: str):
    if PROFILE is None:
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()          # materialise the underlying cudaEvent_t handles
    b.record()
    lib.dta_profile_next_attn(a.cuda_event, b.cuda_event)
    PROFILE.append((kind, a, b))



def _shape(layout: TrieLayout, q, k, v, o, scale: float) -> _cabi.AttnShape:
    s = _cabi.AttnShape()
    s.n_nodes, s.n_pad = layout.n_nodes, layout.n_pad
    s.n_q_heads, s.n_kv_heads, s.head_dim = q.shape[1], k.shape[1], q.shape[2]
    s.q_row_stride, s.q_head_stride = q.stride(0), q.stride(1)
    s.k_row_stride, s.k_head_stride = k.stride(0), k.stride(1)
    s.v_row_stride, s.v_head_stride = v.stride(0), v.stride(1)
    s.o_row_stride, s.o_head_stride = o.stride(0), o.stride(1)
    s.scale = float(scale)
    s.bwd_whole_tiles = int(layout.bwd_whole_tiles(k.shape[1], _sm_count(q.device)))
    return s


_SM_COUNT = {}


def _sm_count(device) -> int:
    idx = torch.device(device).index or 0
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count if torch.cuda.is_available() else 1
    return _SM_COUNT[idx]


def _kernel_ready(t: torch.Tensor) -> torch.Tensor:
    """[N,H,D] bf16 view with unit D stride, 16-byte aligned rows; copy only if needed."""
    ok = (t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and t.data_ptr() % 16 == 0)
    return t if ok else t.contiguous()


def tree_attention_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, layout: TrieLayout, scale: float):
    """Kernel call, no autograd: (q, k, v as handed to the kernel, out [N,Hq,D] bf16, lse [Hq,n_pad] fp32 base 2)."""
    lib = _cabi.load()
    _cabi.require_cuda(q, k, v)
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise _cabi.DtaError("tree attention kernels take bf16 q/k/v")
    if q.shape[0] != layout.n_nodes:
        raise _cabi.DtaError(f"q has {q.shape[0]} rows but the trie has {layout.n_nodes} nodes")
    q, k, v = _kernel_ready(q), _kernel_ready(k), _kernel_ready(v)
    out = torch.empty((q.shape[0], q.shape[1], q.shape[2]), dtype=q.dtype, device=q.device)
    lse = torch.empty((q.shape[1], layout.n_pad), dtype=torch.float32, device=q.device)
    shape = _shape(layout, q, k, v, out, scale)
    meta = layout.meta()
    _arm_profile(lib, "fwd")
    _cabi.check(lib.dta_tree_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                      lse.data_ptr(), C.byref(meta), C.byref(shape),
                                      _cabi.stream_ptr(q.device)))
    return q, k, v, out, lse


def tree_attention_backward(q, k, v, out, lse, dout, layout: TrieLayout, scale: float, dq_fp32: bool = False):
    """Kernel call, no autograd: (dq, dk, dv) with the shapes of q, k, v; dk / dv bf16; dq bf16, or — ``dq_fp32`` —
    the kernel's own fp32 accumulator [n_pad, Hq, D] handed back without the conversion pass (its consumer, the
    q/k-norm + RoPE backward, reads it directly)."""
    lib = _cabi.load()
    dout = dout.contiguous()
    dk = torch.empty((k.shape[0], k.shape[1], k.shape[2]), dtype=k.dtype, device=k.device)
    dv = torch.empty_like(dk)
    if dq_fp32:
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()       # the kernel addresses dk / dv with the k / v strides
        dq = torch.empty((layout.n_pad, q.shape[1], q.shape[2]), dtype=torch.float32, device=q.device)
        shape = _shape(layout, q, dk, dv, out, scale)
    else:
        dq = torch.empty((q.shape[0], q.shape[1], q.shape[2]), dtype=q.dtype, device=q.device)
        # the kernel addresses dq/dk/dv with the q/k/v strides: give it same-stride views
        shape = _shape(layout, dq, dk, dv, out, scale)
        if (q.stride() != dq.stride()) or (k.stride() != dk.stride()) or (v.stride() != dv.stride()):
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    ws = torch.empty(lib.dta_tree_attn_bwd_workspace_bytes(C.byref(shape)), dtype=torch.uint8, device=q.device)
    meta = layout.meta()
    _arm_profile(lib, "bwd")
    fn = lib.dta_tree_attn_bwd_f32dq if dq_fp32 else lib.dta_tree_attn_bwd
    _cabi.check(fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dq.data_ptr(),
                   dk.data_ptr(), dv.data_ptr(), C.byref(meta), C.byref(shape), ws.data_ptr(), _cabi.stream_ptr(q.device)))
    return dq, dk, dv


class TreeAttnFunction(torch.autograd.Function):
    """out[N,Hq,D] = softmax_over_ancestors(q k^T * scale) v   (bf16, fp32 accumulate)."""

    @staticmethod
    def forward(ctx, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, layout: TrieLayout, scale: float):
        q, k, v, out, lse = tree_attention_forward(q, k, v, layout, scale)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.layout = layout
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = tree_attention_backward(q, k, v, out, lse, dout, ctx.layout, ctx.scale)
        return dq, dk, dv, None, None


class TreeAttnFromSaved(torch.autograd.Function):
    """The attention node of a RECOMPUTED layer (fused_layers._RecomputedLayer): the forward kernel already ran in
    the first pass and its (out, lse) were kept, so the node hands `out` back without a launch; its backward is the
    ordinary backward kernel on the recomputed (bit-identical) q, k, v."""

    @staticmethod
    def forward(ctx, q, k, v, layout: TrieLayout, scale: float, out: torch.Tensor, lse: torch.Tensor):
        q, k, v = _kernel_ready(q), _kernel_ready(k), _kernel_ready(v)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.layout = layout
        ctx.scale = scale
        return out.view(out.shape)

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = tree_attention_backward(q, k, v, out, lse, dout, ctx.layout, ctx.scale)
        return dq, dk, dv, None, None, None, None


def tree_attention(q, k, v, layout: TrieLayout, scale: Optional[float] = None) -> torch.Tensor:
    """q [N,Hq,D], k/v [N,Hkv,D] bf16 -> out [N,Hq,D]; differentiable."""
    if scale is None:
        scale = q.shape[-1] ** -0.5
    return TreeAttnFunction.apply(q, k, v, layout, scale)


# ---------------------------------------------------------------------------------------
# HF plugin
# ---------------------------------------------------------------------------------------
@contextmanager
def active_layout(layout: TrieLayout):
    prev = getattr(_tls, "layout", None)
    _tls.layout = layout
    try:
        yield
    finally:
        _tls.layout = prev


def hf_tree_attention_forward(module, query, key, value, attention_mask=None, dropout: float = 0.0,
                              scaling: Optional[float] = None, sliding_window=None, **kwargs):
    """``AttentionInterface`` function: query [1,Hq,N,D], key/value [1,Hkv,N,D] -> ([1,N,Hq,D], None)."""
    layout = kwargs.get("dta_layout") or getattr(_tls, "layout", None)
    if layout is None:
        raise _cabi.DtaError("dta_tree attention called outside a TreeTrainingEngine step (no trie layout)")
    if query.shape[0] != 1:
        raise _cabi.DtaError("dta_tree attention expects batch size 1 (one linearised trie)")
    if dropout:
        raise _cabi.DtaError("attention dropout is not supported by the tree attention kernels")
    if sliding_window is not None:
        raise _cabi.DtaError("sliding-window layers are not supported by the tree attention kernels")
    q = query[0].transpose(0, 1)
    k = key[0].transpose(0, 1)
    v = value[0].transpose(0, 1)
    out = tree_attention(q, k, v, layout, scaling)
    return out.unsqueeze(0), None


_registered = False


def register() -> str:
    """Install the kernels as HF attention implementation ``"dta_tree"`` (idempotent)."""
    global _registered
    if not _registered:
        from transformers import AttentionInterface
        AttentionInterface.register(ATTN_NAME, hf_tree_attention_forward)
        _registered = True
    return ATTN_NAME


@contextmanager
def use_tree_attention(model, layout: TrieLayout):
    """Route every attention layer of ``model`` through the tree kernels for one step."""
    register()
    cfg = model.config
    prev = cfg._attn_implementation
    cfg._attn_implementation = ATTN_NAME
    try:
        with active_layout(layout):
            yield
    finally:
        cfg._attn_implementation = prev
