import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jtransforms_b200 as jt
from oracle import jt_oracle as o
for n in [9216, 10368, 100000, 27000, 12000, 75600, 1562500]:
    x = o.fill_uniform(2 * n, seed=4, lo=-1.0, hi=1.0)
    a = x.copy()
    jt.DoubleFFT_1D(n).complexForward(a)
    want = o.complex_forward_1d(x, n)
    g = a.view(np.complex128); w = want.view(np.complex128)
    bad = np.nonzero(~(np.abs(g - w) < 1e-6))[0]
    print(n, o.rel_l2(a, want), len(bad), bad[:12], bad[-4:] if len(bad) else "", flush=True)
