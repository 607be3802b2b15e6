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
