"""Run in a fresh process with JTB_NO_FAST=1 JTB_NO_FAST2=1 (the lean kernels are switched off when the library is
loaded): with the single-pass limits forced down to 2^5 / 2^3 the general driver needs its three-pass path
(Engine::c2c_big_contig) for 2^11 and 2^12 points."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import jtransforms_b200 as jt
from jtransforms_b200 import _lib
import parity_cases as pc

_lib.use(sys.argv[1])
L = _lib.get()
L.jtb_debug_set_limits(5, 3)
c0 = L.jtb_launch_count(0)
pc.fft1d_complex(jt, "Double", 2048)
assert L.jtb_launch_count(0) - c0 >= 12, "three-pass path not taken"
pc.fft1d_complex(jt, "Double", 4096)
pc.fft1d_complex(jt, "Float", 4096)
pc.fft1d_batch(jt, "Double", 2048, 3, pad=4)
pc.fft1d_real(jt, "Double", 4096)
L.jtb_debug_set_limits(0, 0)
print("three-pass ok")
