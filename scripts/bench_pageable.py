"""Host-array API through pageable vs page-locked caller memory (jtb_host_register): DoubleFFT_3D 256^3 (256 MiB)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jtransforms_b200 as jt
from jtransforms_b200.utils import pinned

n = 256
a = np.random.rand(2 * n ** 3)
f = jt.DoubleFFT_3D(n, n, n)
f.complexForward(a)
def run(reps=5):
    t0 = time.perf_counter()
    for _ in range(reps): f.complexForward(a)
    return (time.perf_counter() - t0) / reps * 1e3
print("pageable numpy array : %.2f ms per transform (%.1f GB/s each way if copy-bound)" % ((ms := run()), 2 * a.nbytes / ms / 1e6))
with pinned(a):
    print("registered (pinned)  : %.2f ms per transform (%.1f GB/s each way if copy-bound)" % ((ms := run()), 2 * a.nbytes / ms / 1e6))
