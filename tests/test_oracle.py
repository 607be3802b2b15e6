"""Pins the CPU oracle (oracle/jt_oracle.py) against the reference's golden
vectors and against literal restatements of the reference's algorithms."""
import os

import numpy as np
import pytest

from oracle import jt_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fftw")
SIZES = [int(s) for s in open(os.path.join(GOLD, "sizes.txt")).read().split()]


def load(n):
    a = np.fromfile(os.path.join(GOLD, f"fftw{n}.in"), dtype="<f8")
    b = np.fromfile(os.path.join(GOLD, f"fftw{n}.out"), dtype="<f8")
    assert a.size == 2 * n and b.size == 2 * n
    return a, b


@pytest.mark.parametrize("n", SIZES)
def test_complex_forward_matches_fftw_golden(n):
    # DoubleFFT_1DTest.testComplexForward (:208-219): RMSE <= 1e-12
    a, want = load(n)
    got = O.complex_forward_1d(a, n)
    assert O.rmse(got, want) <= 1e-12
    assert O.rel_l2(got, want) <= 1e-15 * max(1, np.log2(max(n, 2))) * 4


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 16, 100, 211, 1693])
def test_bluestein_dataflow_is_dft(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert O.rel_l2(O.c2i(O.bluestein_forward_sim(x)), O.c2i(np.fft.fft(x))) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 100, 310])
def test_real_1d_family(n):
    rng = np.random.default_rng(100 + n)
    x = rng.uniform(-1, 1, n)
    full = np.fft.fft(x)
    p = O.real_forward_1d(x, n)
    # testRealForward (:326-364): packed half spectrum vs complexForward
    if n > 1:
        X = O.unpack_real_1d(p, n)
        assert np.allclose(X, full[: n // 2 + 1], atol=1e-12)
    # round trip incl. the n/2-vs-n quirk (DoubleFFT_1DTest.java:573-596)
    back = O.real_inverse_1d(p, n, True)
    assert np.allclose(back, x, atol=1e-12)
    back = O.real_inverse_1d(p, n, False)
    if n > 1:
        f = n / 2.0 if O.is_pow2(n) else n
        assert np.allclose(back, f * x, atol=1e-10)
    a = np.zeros(2 * n)
    a[:n] = x
    assert np.allclose(O.real_forward_full_1d(a, n), O.c2i(full), atol=1e-12)
    assert np.allclose(O.real_inverse_full_1d(a, n, True), O.c2i(np.fft.ifft(x)), atol=1e-12)
    if n > 1:
        q = O.real_inverse2_1d(x, n, False)
        assert np.allclose(O.unpack_real_1d(q, n), np.conj(full[: n // 2 + 1]), atol=1e-12)


@pytest.mark.parametrize("shape", [(2, 2), (4, 4), (8, 8), (4, 16), (16, 4), (2, 8), (8, 2), (32, 8)])
def test_real_2d_layout(shape):
    R, C = shape
    rng = np.random.default_rng(R * 100 + C)
    x = rng.uniform(0, 1, R * C)
    p = O.real_forward_2d(x, R, C)
    # (1) literal restatement of the reference's 3-step algorithm
    assert np.allclose(O.sim_real_forward_2d(x, R, C), p, atol=1e-12)
    # (2) RealFFTUtils_2D.getIndex unpack == realForwardFull (RealFFTUtils_2DTest :69-77)
    full = O.real_forward_full_2d(x, R, C)
    assert np.allclose(O.unpack_real_2d(p, R, C).ravel(), full, atol=1e-12)
    assert np.allclose(O.real_inverse_2d(p, R, C, True), x, atol=1e-12)
    assert np.allclose(O.real_inverse_2d(p, R, C, False), x * R * C / 2, atol=1e-9)


@pytest.mark.parametrize("shape", [(2, 2, 2), (4, 4, 4), (8, 4, 2), (2, 8, 4), (4, 2, 8), (8, 8, 8), (16, 4, 8)])
def test_real_3d_layout(shape):
    S, R, C = shape
    rng = np.random.default_rng(S * 10000 + R * 100 + C)
    x = rng.uniform(0, 1, S * R * C)
    p = O.real_forward_3d(x, S, R, C)
    assert np.allclose(O.sim_real_forward_3d(x, S, R, C), p, atol=1e-12)
    F = O.unpack_real_3d(p, S, R, C)
    assert np.allclose(F, np.fft.fftn(x.reshape(S, R, C)), atol=1e-11)
    assert np.allclose(O.real_inverse_3d(p, S, R, C, True), x, atol=1e-12)


def _dct2_direct(x):
    n = len(x)
    j = np.arange(n)
    return np.array([np.sum(x * np.cos(np.pi * (j + 0.5) * k / n)) for k in range(n)])


def _dct3_direct(a):
    n = len(a)
    k = np.arange(n)
    return np.array([np.sum(a * np.cos(np.pi * k * (j + 0.5) / n)) for j in range(n)])


def _dst2_direct(x):
    n = len(x)
    j = np.arange(n)
    return np.array([np.sum(x * np.sin(np.pi * (j + 0.5) * (k + 1) / n)) for k in range(n)])


@pytest.mark.parametrize("n", [2, 4, 8, 16, 3, 5, 6, 12, 100])
def test_dct_dst_dht_definitions(n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, n)
    f = 1.0 if O.is_pow2(n) else 2.0
    assert np.allclose(O.dct_forward_1d(x, False), f * _dct2_direct(x), atol=1e-12)
    assert np.allclose(O.dst_forward_1d(x, False), f * _dst2_direct(x), atol=1e-12)
    if O.is_pow2(n):
        assert np.allclose(O.dct_inverse_1d(x, False), _dct3_direct(x), atol=1e-12)
    else:
        a = x.copy()
        a[0] *= 0.5
        assert np.allclose(O.dct_inverse_1d(x, False), _dct3_direct(a) / n, atol=1e-12)
    # reference round-trip tests (DoubleDCT_1DTest :125-160 etc.)
    assert np.allclose(O.dct_inverse_1d(O.dct_forward_1d(x, True), True), x, atol=1e-12)
    assert np.allclose(O.dst_inverse_1d(O.dst_forward_1d(x, True), True), x, atol=1e-12)
    assert np.allclose(O.dht_inverse_1d(O.dht_forward_1d(x), True), x, atol=1e-12)
    j = np.arange(n)
    H = np.array([np.sum(x * (np.cos(2 * np.pi * j * k / n) + np.sin(2 * np.pi * j * k / n))) for k in range(n)])
    assert np.allclose(O.dht_forward_1d(x), H, atol=1e-12)


def _nonpow2_dct_forward_sim(x):
    """Literal restatement of the non-pow2 branch, dct/DoubleDCT_1D.java:196-241."""
    n = len(x)
    t = np.concatenate([x, x[::-1]])
    p = O.real_forward_1d(t, 2 * n)
    i = np.arange(n)
    wr = np.cos(np.pi * i / (2 * n))        # makect(n) :523-538: c[2j] = cos(j*pi/2n)
    wi = -np.sin(np.pi * i / (2 * n))       #                     c[2j+1] = -sin(j*pi/2n)
    tr, ti = p[0:2 * n:2], p[1:2 * n:2]      # t[2i], t[2i+1] of the packed 2n-point realForward
    return wr * tr - wi * ti                 # :231-233 (wi[0] = 0, so t[1] = Re[n] is unused)


@pytest.mark.parametrize("n", [3, 5, 6, 12])
def test_dct_nonpow2_branch(n):
    rng = np.random.default_rng(7 * n)
    x = rng.uniform(-1, 1, n)
    assert np.allclose(_nonpow2_dct_forward_sim(x), O.dct_forward_1d(x, False), atol=1e-12)


@pytest.mark.parametrize("shape", [(4, 4), (8, 4), (6, 10), (7, 5), (4, 4, 4), (2, 4, 8), (3, 5, 6), (6, 4, 2)])
def test_dht_nd_equals_ytransform(shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.uniform(0, 1, int(np.prod(shape)))
    assert np.allclose(O.sim_dht_nd(x, shape), O.dht_forward_nd(x, shape), atol=1e-11)


def test_java_random_known_values():
    # java.util.Random(42).nextInt() first value is -1170105035 (well-known)
    r = O.JavaRandom(42)
    v = r._next(32)
    v = v - (1 << 32) if v >= (1 << 31) else v
    assert v == -1170105035
    d = O.JavaRandom(2).doubles(3)
    assert np.all((d >= 0) & (d < 1))


def test_plan_selection():
    assert O.plan_of(1 << 20) == "SPLIT_RADIX"
    assert O.plan_of(1000003) == "BLUESTEIN"
    assert O.plan_of(10158) == "BLUESTEIN"
    assert O.plan_of(65530) == "BLUESTEIN"
    assert O.plan_of(1056) == "MIXED_RADIX"
    assert O.next_pow2(2 * 1000003 - 1) == 1 << 21
