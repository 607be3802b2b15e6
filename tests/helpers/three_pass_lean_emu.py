"""Run in a fresh process with JTB_THREEPASS_MIN=19: the lean three-sweep 1-D path (fast_threepass_contig) at a size
the emulator finishes quickly."""
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
c0 = L.jtb_launch_count(0)
pc.fft1d_complex(jt, "Double", 1 << 19)
assert L.jtb_launch_count(0) - c0 == 9, "expected three launches per transform"
pc.fft1d_complex(jt, "Float", 1 << 20)
print("three-pass lean ok")
