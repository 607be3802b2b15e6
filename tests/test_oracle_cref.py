"""The C restatement of the reference's CPU algorithm (oracle/jt_ref.c, the timed CPU baseline) against the
NumPy oracle and the reference's FFTW golden vectors."""
import os

import numpy as np
import pytest

from oracle import cref, jt_oracle as o

FFTW = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fftw")


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 8192, 16384, 65536, 131072])
def test_cref_fftw_golden(n):
    x = np.fromfile(os.path.join(FFTW, "fftw%d.in" % n), dtype="<f8")
    want = np.fromfile(os.path.join(FFTW, "fftw%d.out" % n), dtype="<f8")
    for nt in (1, 2, 4):
        a = x.copy()
        cref.cfft1d(a, n, -1, nt)
        assert o.rel_l2(a, want) < 1e-12 * max(1, np.log2(n))


def test_cref_inverse_and_nd():
    x = o.fill_uniform(2 * 4096, seed=3)
    a = x.copy()
    cref.cfft1d(a, 4096, +1, 4)
    assert o.rel_l2(a, o.complex_inverse_1d(x, 4096, False)) < 1e-12 * 12
    for dims in [(1, 16, 32), (8, 4, 16), (32, 32, 32), (2, 64, 8)]:
        S, R, Cn = dims
        x = o.fill_uniform(2 * S * R * Cn, seed=4)
        a = x.copy()
        cref.cfft3d(a, S, R, Cn, -1, 3)
        want = o.complex_forward_3d(x, S, R, Cn) if S > 1 else o.complex_forward_2d(x, R, Cn)
        assert o.rel_l2(a, want) < 1e-12 * 20


def test_cref_real2d_r2r_bluestein():
    """the CPU-baseline drivers of configs 2, 3 and 4 compute what the oracle says"""
    for (R, Cn) in [(8, 16), (64, 32), (128, 256)]:
        x = o.fill_uniform(R * Cn, seed=6, lo=-1.0, hi=1.0)
        a = x.copy()
        cref.rfft2d(a, R, Cn, 3)
        assert o.rel_l2(a, o.real_forward_2d(x, R, Cn)) < 1e-12 * 20
        for kind, want in (("dct", o.dct_forward_nd(x, (R, Cn), True)), ("dst", o.dst_forward_nd(x, (R, Cn), True)),
                           ("dht", o.dht_forward_nd(x, (R, Cn)))):
            b = x.copy()
            cref.r2r2d(b, R, Cn, kind, 3)
            assert o.rel_l2(b, want) < 1e-12 * 20, kind
    n, nb = 1009, 3
    x = o.fill_uniform(2 * n * nb, seed=8, lo=-1.0, hi=1.0).astype(np.float32)
    a = x.copy()
    cref.bluestein_f32(a, n, nb, 2)
    want = np.concatenate([o.complex_forward_1d(x[2 * n * i:2 * n * (i + 1)].astype(np.float64), n) for i in range(nb)])
    assert o.rel_l2(a.astype(np.float64), want) < 1e-5 * 10
