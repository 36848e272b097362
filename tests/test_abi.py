"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/gemini_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gemini_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from gemini_b200 import _lib

    names = header_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(_lib.lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.lib.gm_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    import gemini_b200 as gm

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gm.GeminiError) as ei:
        gm.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "gemini_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyref" not in text and "gemini_oracle" not in text and "import oracle" not in text, f


def test_fold_chain_len_is_pure_host():
    from gemini_b200 import _lib

    assert _lib.lib.gm_fr_fold_chain_len(19, 3) == 10 + 5 + 3
    assert _lib.lib.gm_fr_fold_chain_len(16, 4) == 15
    assert _lib.lib.gm_fr_fold_chain_len(0, 4) == 0
