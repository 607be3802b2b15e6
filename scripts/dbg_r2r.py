import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import jtransforms_b200 as jt
from oracle import jt_oracle as o
for prec, dt in (("Double", np.float64), ("Float", np.float32)):
    for n in (512, 1024, 2048, 4096, 8192):
        for kind in ("DCT", "DST", "DHT"):
            x = o.fill_uniform(n, seed=5, lo=-1, hi=1).astype(dt)
            a = x.copy()
            t = getattr(jt, prec + kind + "_1D")(n)
            if kind == "DHT":
                t.forward(a); want = o.dht_forward_nd(x.astype(np.float64), (n,))
            else:
                t.forward(a, True); want = (o.dct_forward_nd if kind == "DCT" else o.dst_forward_nd)(x.astype(np.float64), (n,), True)
            print(prec, n, kind, "%.2e" % o.rel_l2(a, want))
