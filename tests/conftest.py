import os
import sys
import types
from pathlib import Path

import pytest

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout at /root/reference")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "reference" in item.keywords and not REFERENCE.exists():
            item.add_marker(skip_ref)


def load_reference():
    """Make the reference's flat modules importable (with the 3-line areal stub)."""
    if "areal" not in sys.modules:
        areal = types.ModuleType("areal")
        plat = types.ModuleType("areal.platforms")
        plat.is_npu_available = False
        areal.platforms = plat
        sys.modules["areal"] = areal
        sys.modules["areal.platforms"] = plat
    if str(REFERENCE) not in sys.path:
        sys.path.append(str(REFERENCE))      # appended: repo modules keep priority


@pytest.fixture
def reference():
    load_reference()
    import importlib
    mods = {}
    for name in ("trie", "token_trie", "tree_time_model", "data_parallel", "vocab_parallel",
                 "tree_training_engine", "dense"):
        mods[name] = importlib.import_module(name)
    return types.SimpleNamespace(**mods)


@pytest.fixture
def cpu_backend(monkeypatch):
    """Oracle-backed stand-ins for the CUDA ops so engine HOST logic can run without a GPU."""
    import fake_backend
    fake_backend.install(monkeypatch)
    return fake_backend
