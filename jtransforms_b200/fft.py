"""org.jtransforms.fft mirror: DoubleFFT_{1,2,3}D / FloatFFT_{1,2,3}D over libjtb200.

Same method names, argument meaning, in-place semantics and packed layouts as the reference
(fft/DoubleFFT_1D.java, fft/DoubleFFT_2D.java, fft/DoubleFFT_3D.java and their Float twins).  Arrays are
numpy arrays (``double[]`` -> 1-D float64; ``double[][]`` / ``double[][][]`` -> C-contiguous 2-D / 3-D
arrays, which share the flat layout) or CUDA torch tensors for device-resident use.  Java overloads
``f(a)``, ``f(a, offa)``, ``f(a, scale)``, ``f(a, offa, scale)`` are all accepted.
"""
from __future__ import annotations

from . import _lib
from ._plan import Plan


def _args(args, want_scale: bool):
    """Decode the Java overload tails (offa) / (scale) / (offa, scale)."""
    offa, scale = 0, False
    if want_scale:
        if len(args) == 1:
            scale = bool(args[0])
        elif len(args) == 2:
            offa, scale = int(args[0]), bool(args[1])
        else:
            raise TypeError("expected (a, scale) or (a, offa, scale)")
    else:
        if len(args) == 1:
            offa = int(args[0])
        elif len(args) > 1:
            raise TypeError("expected (a) or (a, offa)")
    return offa, scale


class _FFT:
    _prec = _lib.F64

    def __init__(self, *dims, device: int = 0, devices=None):
        self._plan = Plan(_lib.FFT, self._prec, dims, device, devices)

    def setDevices(self, devices):
        """Multi-GPU: spread host-array transforms (3-D: slab decomposition; batches: blocks) over these GPUs --
        the role ConcurrencyUtils.setNumberOfThreads plays for the reference's thread pool."""
        self._plan.set_devices(devices)

    # fft/DoubleFFT_1D.java:243-263, fft/DoubleFFT_2D.java:115-213, fft/DoubleFFT_3D.java:145-325
    def complexForward(self, a, *args):
        offa, _ = _args(args, False)
        self._plan.run(_lib.C2C_FORWARD, a, offa)

    # fft/DoubleFFT_1D.java:362-385
    def complexInverse(self, a, *args):
        offa, scale = _args(args, True)
        self._plan.run(_lib.C2C_INVERSE, a, offa, scale)

    # fft/DoubleFFT_1D.java:524-561, fft/DoubleFFT_2D.java:820-838, fft/DoubleFFT_3D.java:1339-1355
    def realForward(self, a, *args):
        offa, _ = _args(args, False)
        self._plan.run(_lib.R2C_PACKED, a, offa)

    # fft/DoubleFFT_1D.java:678-755
    def realForwardFull(self, a, *args):
        offa, _ = _args(args, False)
        self._plan.run(_lib.R2C_FULL, a, offa)

    # fft/DoubleFFT_1D.java:946-989
    def realInverse(self, a, *args):
        offa, scale = _args(args, True)
        self._plan.run(_lib.C2R_PACKED, a, offa, scale)

    # fft/DoubleFFT_1D.java:1112-1195
    def realInverseFull(self, a, *args):
        offa, scale = _args(args, True)
        self._plan.run(_lib.C2R_FULL, a, offa, scale)


class DoubleFFT_1D(_FFT):
    def __init__(self, n, device: int = 0, devices=None):
        super().__init__(n, device=device, devices=devices)

    # extension used by config 3 (the reference loops over offa, fft/FloatFFT_1D.java:243)
    def complexForwardBatch(self, a, howmany: int, dist: int, offa: int = 0):
        self._plan.run(_lib.C2C_FORWARD, a, offa, False, howmany, dist)


class DoubleFFT_2D(_FFT):
    def __init__(self, rows, columns, device: int = 0, devices=None):
        super().__init__(rows, columns, device=device, devices=devices)


class DoubleFFT_3D(_FFT):
    def __init__(self, slices, rows, columns, device: int = 0, devices=None):
        super().__init__(slices, rows, columns, device=device, devices=devices)


class FloatFFT_1D(DoubleFFT_1D):
    _prec = _lib.F32


class FloatFFT_2D(DoubleFFT_2D):
    _prec = _lib.F32


class FloatFFT_3D(DoubleFFT_3D):
    _prec = _lib.F32
