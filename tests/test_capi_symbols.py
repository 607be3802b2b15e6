"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/jtb200.h declares, and
compute entry points fail loudly (no fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "jtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jtb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from jtransforms_b200 import build as b, _lib
    lib = b.build()
    h = ctypes.CDLL(lib)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(h, n), n
    assert set(_lib.SYMBOLS) == set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from jtransforms_b200 import _lib
    import jtransforms_b200 as jt
    old = _lib._lib
    _lib._lib = None
    try:
        with pytest.raises(_lib.JtbError, match="no CUDA device|CUDA"):
            jt.DoubleFFT_1D(8).complexForward(np.zeros(16))
    finally:
        _lib._lib = old
