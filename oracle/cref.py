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
        _lib.jtref_rfft2d.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int]
        _lib.jtref_r2r2d.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int]
        _lib.jtref_bluestein_f32.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int]
    return _lib


def cfft1d(a: np.ndarray, n: int, isgn: int = -1, nthreads: int = 1):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size >= 2 * n
    assert lib().jtref_cfft1d(a.ctypes.data, n, isgn, nthreads) == 0
    return a


def cfft3d(a: np.ndarray, S: int, R: int, Cn: int, isgn: int = -1, nthreads: int = 1):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size >= 2 * S * R * Cn
    assert lib().jtref_cfft3d(a.ctypes.data, S, R, Cn, isgn, nthreads) == 0
    return a


def rfft2d(a: np.ndarray, R: int, Cn: int, nthreads: int = 1):
    """DoubleFFT_2D.realForward (packed layout), power-of-two sizes"""
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size >= R * Cn
    assert lib().jtref_rfft2d(a.ctypes.data, R, Cn, nthreads) == 0
    return a


def r2r2d(a: np.ndarray, R: int, Cn: int, kind: str, nthreads: int = 1):
    """DoubleDCT_2D / DoubleDST_2D forward(scale=True), DoubleDHT_2D.forward; power-of-two sizes"""
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size >= R * Cn
    assert lib().jtref_r2r2d(a.ctypes.data, R, Cn, {"dct": 0, "dst": 1, "dht": 2}[kind], nthreads) == 0
    return a


def bluestein_f32(a: np.ndarray, n: int, nb: int, nthreads: int = 1):
    """FloatFFT_1D.complexForward through the Bluestein path, nb transforms 2n floats apart"""
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.size >= 2 * n * nb
    assert lib().jtref_bluestein_f32(a.ctypes.data, n, nb, nthreads) == 0
    return a
