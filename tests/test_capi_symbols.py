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
        # the memory and multi-GPU entry points refuse as well (JTB_ERR_CUDA = 3): nothing runs on the host
        lib = _lib.get()
        a = np.zeros(64)
        p = ctypes.c_void_p()
        assert lib.jtb_host_alloc(ctypes.byref(p), 64) == _lib.ERR_CUDA
        assert lib.jtb_host_register(ctypes.c_void_p(a.ctypes.data), a.nbytes) == _lib.ERR_CUDA
        arr = (ctypes.c_void_p * 2)(a.ctypes.data, a.ctypes.data)
        assert lib.jtb_fft3d_k1_scatter(0, 0, ctypes.c_void_p(a.ctypes.data), 64, 1, 8, 2, 0, arr, 1, None) == _lib.ERR_CUDA
        assert lib.jtb_lines_c2c_device(0, 0, ctypes.c_void_p(a.ctypes.data), 8, 1, 1, 0, 8, 1, 0, 1.0, None) == _lib.ERR_CUDA
    finally:
        _lib._lib = old
