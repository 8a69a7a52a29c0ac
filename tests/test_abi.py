"""The C-ABI library loads and exports every symbol include/mp2.h declares (no GPU needed, no
compute call made), and compute calls fail loudly -- never fall back -- without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mp2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mp2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from magpurify2_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table covers the header exactly


def test_library_basics():
    from magpurify2_b200 import _lib

    lib = _lib.load()
    assert b"sm_100a" in lib.mp2_version()
    lut = np.zeros(256, dtype=np.uint8)
    assert lib.mp2_canonical_lut(lut.ctypes.data) == 0
    import hashlib

    want = open(os.path.join(ROOT, "tests", "golden", "lut_sha256.txt")).read().strip()
    assert hashlib.sha256(lut.tobytes()).hexdigest() == want


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must raise (MP2_ENODEV), not compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from magpurify2_b200 import _composition, _lib

    with pytest.raises(_lib.Mp2Error) as ei:
        _composition.get_tnf(["ACGTACGT"])
    assert ei.value.code == _lib.ENODEV
    with pytest.raises(_lib.Mp2Error):
        _composition.get_gc(["ACGT"])


def test_product_never_imports_oracle():
    bad = []
    for dp, _dn, fns in os.walk(os.path.join(ROOT, "magpurify2_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                if (re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "libmp2_oracle" in txt
                        or re.search(r"#include\s+[<\"][^>\"]*oracle", txt) or "mp2o_" in txt):
                    bad.append(fn)
    assert not bad, bad
