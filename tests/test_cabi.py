"""The C-ABI library loads without a GPU and exports every symbol include/forte2_b200.h declares."""
import ctypes
import os
import re

import pytest

from forte2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "forte2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(f2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in forte2_b200/_lib.py"
    for name in _lib.SIGNATURES:
        assert name in names, f"{name} bound in Python but not declared in the header"
    assert L.f2_abi_version() == 1


def test_builder_fails_loudly_without_gpu():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    import forte2_b200 as f2
    s = f2.CIStrings(2, 2, 0, [[0] * 4], [], [])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        f2.CISigmaBuilder(s, 0.0, np.zeros((4, 4)), np.zeros((4,) * 4))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "forte2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".hpp", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle/|liboracle|libforte2_ref", src, re.M), \
                    f"{f} reaches into oracle/"
