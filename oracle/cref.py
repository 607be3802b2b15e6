"""ctypes access to oracle/jt_ref.c (TEST/BENCH INFRASTRUCTURE ONLY)."""
import ctypes as C

import numpy as np

from .build_oracle import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.jtref_cfft1d.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int]
        _lib.jtref_cfft3d.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int]
    return _lib


def cfft1d(a: np.ndarray, n: int, isgn: int = -1, nthreads: int = 1):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size >= 2 * n
    assert lib().jtref_cfft1d(a.ctypes.data, n, isgn, nthreads) == 0
    return a


def cfft3d(a: np.ndarray, S: int, R: int, Cn: int, isgn: int = -1, nthreads: int = 1):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size >= 2 * S * R * Cn
    assert lib().jtref_cfft3d(a.ctypes.data, S, R, Cn, isgn, nthreads) == 0
    return a
