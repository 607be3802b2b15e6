"""Subprocess helper: multi-device plan (virtual ranks) with the column-block pipelined exchange switched on
(JTB_SLAB_CHUNKS is read once per process), against the oracle, through the emulated library."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from jtransforms_b200 import _lib  # noqa: E402

_lib.use(sys.argv[1])
import jtransforms_b200 as jt  # noqa: E402
import parity_cases as pc  # noqa: E402

for prec, dims, P in (("Double", (64, 64, 64), 2), ("Double", (64, 64, 32), 4), ("Float", (64, 64, 64), 2)):
    pc.fft3d_multi(jt, prec, dims, [0] * P)
print("ok")
