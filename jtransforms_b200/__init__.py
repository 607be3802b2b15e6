"""jtransforms_b200 -- B200-native drop-in for the JTransforms transform hot path.

Host-side mirror of the org.jtransforms API over the C-ABI CUDA library libjtb200.so
(include/jtb200.h).  No CPU fallback exists: without the built library and a CUDA device every
transform raises.
"""
from . import _lib  # noqa: F401
from .fft import (DoubleFFT_1D, DoubleFFT_2D, DoubleFFT_3D, FloatFFT_1D, FloatFFT_2D, FloatFFT_3D)  # noqa: F401
from .realfftutils import RealFFTUtils_2D, RealFFTUtils_3D  # noqa: F401
from .utils import CommonUtils, ConcurrencyUtils  # noqa: F401
from .r2r import *  # noqa: F401,F403
from .r2r import (DoubleDCT_1D, DoubleDCT_2D, DoubleDCT_3D, DoubleDST_1D, DoubleDST_2D, DoubleDST_3D,  # noqa: F401
                  DoubleDHT_1D, DoubleDHT_2D, DoubleDHT_3D, FloatDCT_1D, FloatDCT_2D, FloatDCT_3D,
                  FloatDST_1D, FloatDST_2D, FloatDST_3D, FloatDHT_1D, FloatDHT_2D, FloatDHT_3D)

__version__ = "0.1.0"
