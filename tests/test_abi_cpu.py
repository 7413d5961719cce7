"""CPU tests of the boundary: the C-ABI library loads, exports every symbol
include/gomc_b200.h declares, and refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest

from gomc_b200 import engine as eng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gomc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gomcb200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(eng.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(eng.LIB_PATH):
        pytest.fail("gomc_b200/libgomc_b200.so missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(eng.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert eng.load_library().gomcb200_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(eng.EngineError) as ei:
        eng.Engine(1)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_path_does_not_import_oracle():
    """Nothing under gomc_b200/ may reference oracle/ (checker stays a checker)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gomc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "gomc_oracle" not in txt, f
                assert "from oracle" not in txt and "import oracle" not in txt, f
