"""org.jtransforms.{dct,dst,dht} mirror: Double/Float DCT, DST and DHT in 1/2/3 dimensions.

dct/DoubleDCT_1D.java:169-243 (forward), :361-434 (inverse); dst/DoubleDST_1D.java:96-160, :264-325;
dht/DoubleDHT_1D.java:94-152, :255-294 and the 2-D / 3-D classes beside them.
"""
from __future__ import annotations

from . import _lib
from ._plan import Plan
from .fft import _args


class _R2R:
    _prec = _lib.F64
    _kind = _lib.DCT

    def __init__(self, *dims, device: int = 0):
        self._plan = Plan(self._kind, self._prec, dims, device)

    def forward(self, a, *args):
        offa, scale = _args(args, True)
        self._plan.run(_lib.R2R_FORWARD, a, offa, scale)

    def inverse(self, a, *args):
        offa, scale = _args(args, True)
        self._plan.run(_lib.R2R_INVERSE, a, offa, scale)


class _DHT(_R2R):
    _kind = _lib.DHT

    # DoubleDHT_*.forward takes no scale argument (dht/DoubleDHT_1D.java:94)
    def forward(self, a, *args):
        offa, _ = _args(args, False)
        self._plan.run(_lib.R2R_FORWARD, a, offa, False)


def _make(name, base, kind, prec, rank):
    def __init__(self, *dims, device: int = 0):
        if len(dims) != rank:
            raise TypeError("%s takes %d size argument(s)" % (name, rank))
        base.__init__(self, *dims, device=device)
    return type(name, (base,), {"_kind": kind, "_prec": prec, "__init__": __init__})


for _p, _pn in ((_lib.F64, "Double"), (_lib.F32, "Float")):
    for _r in (1, 2, 3):
        globals()["%sDCT_%dD" % (_pn, _r)] = _make("%sDCT_%dD" % (_pn, _r), _R2R, _lib.DCT, _p, _r)
        globals()["%sDST_%dD" % (_pn, _r)] = _make("%sDST_%dD" % (_pn, _r), _R2R, _lib.DST, _p, _r)
        globals()["%sDHT_%dD" % (_pn, _r)] = _make("%sDHT_%dD" % (_pn, _r), _DHT, _lib.DHT, _p, _r)
