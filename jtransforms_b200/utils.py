"""org.jtransforms.utils.CommonUtils / pl.edu.icm.jlargearrays.ConcurrencyUtils surface that callers of the hot path
touch.  The arithmetic helpers keep their meaning (utils/CommonUtils.java:215-332); the thread-pool knobs
(utils/CommonUtils.java:47-206) are accepted and ignored -- dispatch is a GPU launch, not a pool."""
from __future__ import annotations


class CommonUtils:
    _thresholds = {"1D_FFT_2Threads": 8192, "1D_FFT_4Threads": 65536, "2D": 65536, "3D": 65536}
    _use_large_arrays = False

    @staticmethod
    def nextPow2(x: int) -> int:
        if x < 1:
            raise ValueError("x must be greater or equal 1")
        return 1 << (int(x) - 1).bit_length()

    @staticmethod
    def prevPow2(x: int) -> int:
        if x < 1:
            raise ValueError("x must be greater or equal 1")
        return 1 << (int(x).bit_length() - 1)

    @staticmethod
    def isPowerOf2(x: int) -> bool:
        return x > 0 and (x & (x - 1)) == 0

    @staticmethod
    def getReminder(n: int, factors) -> int:
        if n <= 0:
            raise ValueError("n must be positive integer")
        rem = int(n)
        for f in factors:
            while rem > 1 and rem % f == 0:
                rem //= f
        return rem

    # thread-begin thresholds: stored so that getters round-trip, otherwise without effect
    @classmethod
    def setThreadsBeginN_1D_FFT_2Threads(cls, n):
        cls._thresholds["1D_FFT_2Threads"] = max(1024, int(n))

    @classmethod
    def setThreadsBeginN_1D_FFT_4Threads(cls, n):
        cls._thresholds["1D_FFT_4Threads"] = max(1024, int(n))

    @classmethod
    def setThreadsBeginN_2D(cls, n):
        cls._thresholds["2D"] = max(4096, int(n))

    @classmethod
    def setThreadsBeginN_3D(cls, n):
        cls._thresholds["3D"] = max(4096, int(n))

    @classmethod
    def getThreadsBeginN_1D_FFT_2Threads(cls):
        return cls._thresholds["1D_FFT_2Threads"]

    @classmethod
    def getThreadsBeginN_1D_FFT_4Threads(cls):
        return cls._thresholds["1D_FFT_4Threads"]

    @classmethod
    def getThreadsBeginN_2D(cls):
        return cls._thresholds["2D"]

    @classmethod
    def getThreadsBeginN_3D(cls):
        return cls._thresholds["3D"]

    @classmethod
    def resetThreadsBeginN(cls):
        cls._thresholds.update({"2D": 65536, "3D": 65536})

    @classmethod
    def resetThreadsBeginN_FFT(cls):
        cls._thresholds.update({"1D_FFT_2Threads": 8192, "1D_FFT_4Threads": 65536})

    @classmethod
    def setUseLargeArrays(cls, flag: bool):
        cls._use_large_arrays = bool(flag)

    @classmethod
    def isUseLargeArrays(cls) -> bool:
        return cls._use_large_arrays


class ConcurrencyUtils:
    """number-of-threads knob of the external JLargeArrays pool: kept for source compatibility, no effect"""
    _n = 1

    @classmethod
    def setNumberOfThreads(cls, n: int):
        cls._n = max(1, int(n))

    @classmethod
    def getNumberOfThreads(cls) -> int:
        return cls._n


class pinned:
    """Context manager that page-locks a numpy array for the duration of a block (jtb_host_register), so the
    host-array API copies it at DMA speed -- what a Java caller does once for an off-heap segment / LargeArray::

        with pinned(a):
            fft.complexForward(a)
    """

    def __init__(self, a):
        self.a = a

    def __enter__(self):
        import ctypes as C
        from . import _lib
        _lib.check(_lib.get().jtb_host_register(C.c_void_p(self.a.ctypes.data), self.a.nbytes))
        return self.a

    def __exit__(self, *exc):
        import ctypes as C
        from . import _lib
        _lib.check(_lib.get().jtb_host_unregister(C.c_void_p(self.a.ctypes.data)))
        return False
