import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jtransforms_b200 as jt
from oracle import jt_oracle as o
n = 4096
x = o.fill_uniform(n, seed=5, lo=-1, hi=1)
for rep in range(2):
    a = x.copy()
    jt.DoubleDCT_1D(n).forward(a, True)
    want = o.dct_forward_nd(x, (n,), True)
    bad = np.nonzero(np.abs(a - want) > 1e-9)[0]
    print("rep", rep, "bad count", len(bad), "first", bad[:20], "last", bad[-10:])
    if len(bad):
        print(a[bad[:6]], want[bad[:6]])
import scipy.fft as sfft
xr = sfft.idct(a, type=2, norm="ortho")
d = np.abs(xr - x)
badi = np.nonzero(d > 1e-9)[0]
print("input-domain differences:", len(badi), badi[:40], badi[-10:])
print(xr[badi[:8]], x[badi[:8]])
