"""RealFFTUtils_2D/3D (host-side layout arithmetic) against the oracle's packed layouts -- the reference's own
check is unpack(getIndex) == realForwardFull (src/test/java/org/jtransforms/fft/RealFFTUtils_2DTest.java:69-77)."""
import numpy as np
import pytest

from jtransforms_b200.realfftutils import MIN_VALUE, RealFFTUtils_2D, RealFFTUtils_3D
from oracle import jt_oracle as o


@pytest.mark.parametrize("dims", [(2, 2), (4, 8), (8, 4), (16, 16), (2, 16), (32, 8)])
def test_realfftutils_2d(dims):
    R, C = dims
    x = o.fill_uniform(R * C, seed=5)
    packed = o.real_forward_2d(x, R, C)
    full = o.real_forward_full_2d(x, R, C).reshape(R, 2 * C)
    u = RealFFTUtils_2D(R, C)
    for r in range(R):
        for c in range(2 * C):
            assert abs(u.unpack(r, c, packed) - full[r, c]) < 1e-9
            want = o.real2d_get_index(R, C, r, c)       # literal restatement of the reference's ladder
            got = u.getIndex(r, c)
            assert (got == MIN_VALUE) == (want is None)
            if want is not None:
                assert got == want
    q = np.zeros_like(packed)
    for r in range(R):
        for c in range(2 * C):
            if u.getIndex(r, c) != MIN_VALUE:
                u.pack(full[r, c], r, c, q)
    assert np.allclose(q, packed)


@pytest.mark.parametrize("dims", [(2, 2, 2), (4, 4, 8), (8, 2, 4), (2, 8, 4), (8, 8, 8)])
def test_realfftutils_3d(dims):
    S, R, C = dims
    x = o.fill_uniform(S * R * C, seed=6)
    packed = o.real_forward_3d(x, S, R, C)
    full = o.real_forward_full_3d(x, S, R, C).reshape(S, R, 2 * C)
    u = RealFFTUtils_3D(S, R, C)
    for s in range(S):
        for r in range(R):
            for c in range(2 * C):
                assert abs(u.unpack(s, r, c, packed) - full[s, r, c]) < 1e-9


def test_common_utils():
    from jtransforms_b200 import CommonUtils, ConcurrencyUtils
    assert CommonUtils.nextPow2(1) == 1 and CommonUtils.nextPow2(5) == 8 and CommonUtils.nextPow2(1024) == 1024
    assert CommonUtils.prevPow2(5) == 4
    assert CommonUtils.isPowerOf2(64) and not CommonUtils.isPowerOf2(96) and not CommonUtils.isPowerOf2(0)
    assert CommonUtils.getReminder(1000003, (4, 2, 3, 5)) == 1000003      # prime -> Bluestein in the reference
    assert CommonUtils.getReminder(120, (4, 2, 3, 5)) == 1
    with pytest.raises(ValueError):
        CommonUtils.nextPow2(0)
    CommonUtils.setThreadsBeginN_2D(4)
    assert CommonUtils.getThreadsBeginN_2D() == 4096                       # floor, utils/CommonUtils.java:152-160
    ConcurrencyUtils.setNumberOfThreads(8)
    assert ConcurrencyUtils.getNumberOfThreads() == 8
