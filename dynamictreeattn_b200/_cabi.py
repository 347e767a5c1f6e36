"""ctypes binding of libdta_b200.so — the ONLY route from Python to the kernels.

Signatures mirror ``include/dta_b200.h`` one to one.  There is deliberately no
fallback: if the shared library is missing or fails to load, every device-side
op raises.  PyTorch is used for memory and streams only; the library sees raw
device pointers and a ``cudaStream_t``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import torch

# DTA_LIB_PATH: load another build of the same ABI (A/B timing of kernel changes on one box); default: the in-tree build
_LIB_PATH = Path(os.environ.get("DTA_LIB_PATH") or Path(__file__).resolve().parent / "lib" / "libdta_b200.so")
_lib = None

ABI_VERSION = 8
TILE = 128


class DtaError(RuntimeError):
    pass


class AttnShape(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("n_pad", C.c_int32),
        ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("q_row_stride", C.c_int64), ("q_head_stride", C.c_int64),
        ("k_row_stride", C.c_int64), ("k_head_stride", C.c_int64),
        ("v_row_stride", C.c_int64), ("v_head_stride", C.c_int64),
        ("o_row_stride", C.c_int64), ("o_head_stride", C.c_int64),
        ("scale", C.c_float),
        ("bwd_whole_tiles", C.c_int32),
    ]


class AttnMeta(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("subtree_end", "kmax_end", "kmin_end", "q_order", "k_order", "k_qlast")]


_P, _I, _L, _F, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/dta_b200.h
SIGNATURES = {
    "dta_last_error": (C.c_char_p, []),
    "dta_abi_version": (_I, []),
    "dta_lcp_adjacent": (_I, [_P, _P, _I, _P, _P, _P]),
    "dta_trie_workspace_bytes": (_Z, [_I]),
    "dta_trie_linearize": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "dta_trie_paths": (_I, [_P, _P, _P, _I, _P, _L, _P, _P, _P]),
    "dta_tile_meta": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dta_trie_build_workspace_bytes": (_Z, [_I]),
    "dta_trie_build": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "dta_host_subtrie_stats": (_I, [_P, _P, _I, _I, _L, _P]),
    "dta_tree_attn_fwd": (_I, [_P, _P, _P, _P, _P, C.POINTER(AttnMeta), C.POINTER(AttnShape), _P]),
    "dta_profile_next_attn": (_I, [_P, _P]),
    "dta_tree_attn_bwd_workspace_bytes": (_Z, [C.POINTER(AttnShape)]),
    "dta_tree_attn_bwd": (_I, [_P] * 9 + [C.POINTER(AttnMeta), C.POINTER(AttnShape), _P, _P]),
    "dta_tree_attn_bwd_f32dq": (_I, [_P] * 9 + [C.POINTER(AttnMeta), C.POINTER(AttnShape), _P, _P]),
    "dta_logprob_row_stats": (_I, [_P, _I, _L, _L, _L, _F, _P, _L, _P, _P]),
    "dta_logprob_gather_labels": (_I, [_P, _I, _L, _L, _F, _P, _P, _L, _L, _P, _L, _P]),
    "dta_logprob_finalize": (_I, [_P, _I, _L, _L, _P, _P, _P, _P]),
    "dta_logprob_bwd": (_I, [_P, _I, _L, _L, _L, _F, _P, _P, _P, _P, _P, _L, _P, _L, _P]),
    "dta_logprob_bwd_scatter": (_I, [_P, _I, _L, _L, _F, _P, _P, _P, _L, _L, _P]),
    "dta_logprob_rowsum_grad": (_I, [_P, _P, _L, _P, _P]),
    "dta_lm_head_stats_splits": (_I, [_L, _L]),
    "dta_lm_head_stats_rows_pad": (_L, [_L]),
    "dta_lm_head_stats": (_I, [_P, _L, _P, _L, _L, _L, _I, _F, _P, _P]),
    "dta_lm_head_stats_logits": (_I, [_P, _L, _P, _L, _L, _L, _I, _F, _P, _P, _L, _P]),
    "dta_lm_head_dlogits": (_I, [_P, _L, _P, _L, _L, _L, _I, _F, _P, _P, _P, _P, _P, _L, _P]),
    "dta_lm_head_label_logits": (_I, [_P, _L, _P, _L, _I, _L, _P, _P, _L, _F, _P, _P]),
    "dta_rmsnorm_fwd": (_I, [_P, _P, _P, _L, _I, _F, _P, _P, _P, _P]),
    "dta_rmsnorm_bwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P, _P]),
    "dta_qknorm_rope_fwd": (_I, [_P, _L, _L, _P, _L, _L, _P, _P, _P, _P, _L, _I, _I, _I, _F, _P, _P, _P, _P]),
    "dta_qknorm_rope_bwd": (_I, [_P, _P, _P, _L, _L, _P, _L, _L, _P, _P, _P, _P, _P, _L, _I, _I, _I,
                                 _P, _L, _L, _P, _L, _L, _P, _P, _P]),
    "dta_qknorm_rope_bwd_f32dq": (_I, [_P, _P, _P, _L, _L, _P, _L, _L, _P, _P, _P, _P, _P, _L, _I, _I, _I,
                                       _P, _L, _L, _P, _L, _L, _P, _P, _P]),
    "dta_swiglu_fwd": (_I, [_P, _P, _L, _P, _P]),
    "dta_swiglu_bwd": (_I, [_P, _P, _P, _L, _P, _P, _P]),
}


# kernels launched per successful call (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {
    "dta_lcp_adjacent": 1, "dta_trie_linearize": 2, "dta_trie_paths": 1, "dta_tile_meta": 2, "dta_trie_build": 3,
    "dta_tree_attn_fwd": 1, "dta_tree_attn_bwd": 4, "dta_tree_attn_bwd_f32dq": 4,
    "dta_logprob_row_stats": 1, "dta_logprob_gather_labels": 1, "dta_logprob_finalize": 1,
    "dta_logprob_bwd": 1, "dta_logprob_bwd_scatter": 1, "dta_logprob_rowsum_grad": 1,
    "dta_lm_head_stats": 1, "dta_lm_head_stats_logits": 1, "dta_lm_head_dlogits": 1, "dta_lm_head_label_logits": 1,
    "dta_rmsnorm_fwd": 1, "dta_rmsnorm_bwd": 1, "dta_qknorm_rope_fwd": 1, "dta_qknorm_rope_bwd": 1, "dta_qknorm_rope_bwd_f32dq": 1,
    "dta_swiglu_fwd": 1, "dta_swiglu_bwd": 1,
}
launch_count = 0          # kernels of OURS launched by this process so far
call_count: dict = {}


class _Counted:
    """Callable proxy over one library entry point that keeps the launch tally."""
    __slots__ = ("fn", "name", "n")

    def __init__(self, fn, name, n):
        self.fn, self.name, self.n = fn, name, n

    def __call__(self, *args):
        global launch_count
        rc = self.fn(*args)
        if rc == 0:
            launch_count += self.n
            call_count[self.name] = call_count.get(self.name, 0) + 1
        return rc


class _Lib:
    def __init__(self, cdll):
        self._cdll = cdll
        for name in SIGNATURES:
            fn = getattr(cdll, name)
            n = KERNELS_PER_CALL.get(name)
            setattr(self, name, _Counted(fn, name, n) if n else fn)


def lib_path() -> Path:
    return _LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if allowed and needed) the C-ABI library; raise if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    in_tree = "DTA_LIB_PATH" not in os.environ
    may_build = build_if_missing and os.environ.get("DTA_NO_BUILD") != "1"
    if not _LIB_PATH.exists():
        if not may_build or not in_tree:
            raise DtaError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        from ._build import build_library
        build_library()
    elif in_tree:
        # a library built from OTHER sources than the ones in the tree must not load silently (same ABI number,
        # different kernels): compare the build stamp with a fresh digest of csrc/ + the header + the flags
        from ._build import LIBDIR, _digest, build_library
        stamp = LIBDIR / "build.sha256"
        if not stamp.exists() or stamp.read_text().strip() != _digest():
            if not may_build:
                raise DtaError(f"{_LIB_PATH} is stale (csrc/ or include/dta_b200.h changed since it was built): run "
                               "`python -c 'import __graft_entry__ as g; g.build()'`")
            build_library()
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.dta_abi_version() != ABI_VERSION:
        raise DtaError(f"ABI mismatch: library {lib.dta_abi_version()} vs binding {ABI_VERSION}")
    _lib = _Lib(lib)
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise DtaError(f"libdta_b200 error {rc}: {load().dta_last_error().decode(errors='replace')}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DtaError("dynamictreeattn_b200 device ops need CUDA tensors (there is no CPU fallback)")
