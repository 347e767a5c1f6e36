"""The C-ABI library loads and exports every symbol include/dta_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "dta_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dta_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from dynamictreeattn_b200 import _cabi
    lib = _cabi.load()
    names = _declared()
    assert len(names) >= 17
    raw = ctypes.CDLL(str(_cabi.lib_path()))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_cabi.SIGNATURES) == names
    assert lib.dta_abi_version() == _cabi.ABI_VERSION
    assert lib.dta_trie_workspace_bytes(1000) >= 2 * 11 * 1000 * 4


def test_argument_validation_without_gpu():
    """Invalid sizes are rejected on the host before any launch; the message is retrievable."""
    from dynamictreeattn_b200 import _cabi
    lib = _cabi.load()
    rc = lib.dta_trie_linearize(None, None, 0, 0, 0, None, None, None, None, None, None, None)
    assert rc != 0
    assert b"n_leaf" in lib.dta_last_error()
    rc = lib.dta_logprob_row_stats(None, 7, 1, 1, 1, 1.0, None, 0, None, None)
    assert rc != 0 and b"dtype" in lib.dta_last_error()


def test_no_cpu_fallback():
    import torch
    from dynamictreeattn_b200 import _cabi
    from dynamictreeattn_b200.attention import tree_attention
    from dynamictreeattn_b200.linearize import linearize
    from dynamictreeattn_b200.vocab_parallel import gather_logprobs
    with pytest.raises(_cabi.DtaError):
        linearize([torch.tensor([1, 2, 3])], [3], [], "cpu")
    with pytest.raises(_cabi.DtaError):
        gather_logprobs(torch.randn(4, 10), torch.zeros(4, dtype=torch.long))
    with pytest.raises(_cabi.DtaError):
        q = torch.randn(4, 2, 128).bfloat16()
        tree_attention(q, q[:, :1], q[:, :1], None)


def test_product_never_imports_oracle():
    pkg = ROOT / "dynamictreeattn_b200"
    for p in pkg.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M), p


def test_bench_dense_leg_records_failure_instead_of_raising():
    """bench.py's dense-causal comparison leg runs after the primary numbers are taken and must never take the bench
    line down: without a GPU the product log-prob op refuses to run (no CPU fallback) and the leg reports that."""
    import sys
    from pathlib import Path
    import torch
    from transformers import Qwen3Config, Qwen3ForCausalLM
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench
    if torch.cuda.is_available():
        import pytest
        pytest.skip("failure path is the CPU-only one")
    cfg = Qwen3Config(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                      num_key_value_heads=1, head_dim=16)
    model = Qwen3ForCausalLM(cfg)
    seqs = [torch.randint(0, 64, (12,)), torch.randint(0, 64, (9,))]
    out = bench.dense_gpu_leg(model, seqs, [dict(bench.ATTACH) for _ in seqs], torch.device("cpu"))
    assert out["value"] is None and "error" in out and out["unit"] == "sequence tokens/s"
    assert all(p.grad is None for p in model.parameters())


def test_stale_library_is_refused(monkeypatch):
    """_cabi.load() compares lib/build.sha256 with a fresh digest of the sources: a library of the same ABI number
    built from other sources raises instead of loading silently (when building is not allowed)."""
    from dynamictreeattn_b200 import _build, _cabi
    _cabi.load()                                        # make sure a current build exists
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_build, "_digest", lambda: "0" * 64)
    with pytest.raises(_cabi.DtaError, match="stale"):
        _cabi.load(build_if_missing=False)
    monkeypatch.undo()
    _cabi._lib = None
    assert _cabi.load(build_if_missing=False) is not None


def test_header_is_plain_c_and_links(tmp_path):
    """include/dta_b200.h is the boundary a non-Python host binds: it must compile as strict C99 (no C++ or torch
    types in any signature) and a C program must link against the in-tree library and read its ABI version."""
    import shutil
    import subprocess
    from dynamictreeattn_b200 import _cabi
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    _cabi.load()
    src = tmp_path / "abi_test.c"
    src.write_text('#include "dta_b200.h"\n'
                   'int main(void) { return dta_abi_version() == DTA_ABI_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "abi_test"
    lib_dir = _cabi.lib_path().parent
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), str(src),
                        "-o", str(exe), "-L", str(lib_dir), "-ldta_b200", f"-Wl,-rpath,{lib_dir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0
