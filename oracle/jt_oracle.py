"""CPU oracle for the JTransforms transform hot path (TEST INFRASTRUCTURE ONLY).

This module restates, in NumPy/SciPy FP64, *what* every public transform of
wendykierp/JTransforms computes and *where* it leaves each number (packed
layouts, scaling quirks).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / reference arm may import it; the product package
``jtransforms_b200`` never does (it fails loudly without its CUDA library).

Pinning status
--------------
* complexForward (1-D) is pinned against the reference's own FFTW golden
  vectors (``src/test/resources/fftw{n}.in/.out``, consumed by
  ``src/test/java/org/jtransforms/fft/DoubleFFT_1DTest.java:208-219``), copied
  to ``tests/golden/fftw``.  ``tests/test_oracle.py`` checks all 32 sizes.
* Every other FFT-family method is pinned the way the reference's own tests pin
  it: relative to complexForward (DoubleFFT_1DTest.java:243-596,
  DoubleFFT_2DTest.java:167-185,399-476, DoubleFFT_3DTest.java:165-199,
  RealFFTUtils_2DTest.java:69-77) and, for the packed 2-D/3-D layouts, by
  re-running the reference's three-step algorithm (row real FFT, column complex
  FFT incl. the pseudo column, ``rdft2d_sub``/``rdft3d_sub``) step by step in
  NumPy (functions ``sim_*`` below) and by the executable layout specification
  ``RealFFTUtils_2D.getIndex`` (restated in ``real2d_get_index``).
* DCT / DST / DHT absolute values: **parity unpinned** by the reference (its
  tests are round trips only, e.g. DoubleDCT_1DTest.java:125-160).  The oracle
  follows the code's documented definitions (Ooura ``ddct``) and the explicit
  non-power-of-two branches, cross-checked against an O(n^2) direct sum.
* The Java toolchain is absent in this image, so the reference itself cannot be
  executed; see DESIGN.md.

All file:line citations are relative to
``/root/reference/src/main/java/org/jtransforms``.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft


def _workers(x) -> int:
    """Host threads for the big configurations (512^3, 8192^2): same pocketfft arithmetic, just not on one core."""
    import os
    return (os.cpu_count() or 1) if np.size(x) >= (1 << 22) else 1


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------


def is_pow2(n: int) -> bool:
    """utils/CommonUtils.java:254-262 (isPowerOf2)."""
    return n > 0 and (n & (n - 1)) == 0


def next_pow2(n: int) -> int:
    """utils/CommonUtils.java:215-232 (nextPow2)."""
    if n < 1:
        raise ValueError("x must be greater or equal 1")
    return 1 << (n - 1).bit_length()


def get_reminder(n: int, factors=(4, 2, 3, 5)) -> int:
    """utils/CommonUtils.java:317-332 (getReminder): strip the given factors."""
    rem = n
    if n <= 0:
        raise ValueError("n must be positive integer")
    for f in factors:
        while rem > 1 and rem % f == 0:
            rem //= f
    return rem


def plan_of(n: int) -> str:
    """fft/DoubleFFT_1D.java:117-146: which plan the reference picks."""
    if is_pow2(n):
        return "SPLIT_RADIX"
    if get_reminder(n) >= 211:
        return "BLUESTEIN"
    return "MIXED_RADIX"


def c2i(z: np.ndarray) -> np.ndarray:
    """complex array -> interleaved (re, im) float64 array (last axis doubled)."""
    z = np.asarray(z, dtype=np.complex128)
    out = np.empty(z.shape[:-1] + (2 * z.shape[-1],), dtype=np.float64)
    out[..., 0::2] = z.real
    out[..., 1::2] = z.imag
    return out


def i2c(a: np.ndarray) -> np.ndarray:
    """interleaved (re, im) -> complex128."""
    a = np.asarray(a, dtype=np.float64)
    return a[..., 0::2] + 1j * a[..., 1::2]


def rel_l2(got, want) -> float:
    """||got-want||_2 / ||want||_2 (north-star parity metric)."""
    got = np.asarray(got, dtype=np.float64).ravel()
    want = np.asarray(want, dtype=np.float64).ravel()
    den = np.linalg.norm(want)
    num = np.linalg.norm(got - want)
    return float(num / den) if den > 0 else float(num)


def rmse(a, b) -> float:
    """utils/IOUtils.java:185-201 (computeRMSE): sqrt(mean((a-b)^2))."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.sqrt(np.mean((a - b) ** 2)))


class JavaRandom:
    """java.util.Random (Java SE specification LCG).

    The reference seeds it with 20110602 in its tests
    (src/test/java/org/jtransforms/fft/DoubleFFT_1DTest.java:248) and with 2
    in its benchmarks (utils/IOUtils.java:277-283).
    """

    _M = (1 << 48) - 1

    def __init__(self, seed: int):
        self.s = (seed ^ 0x5DEECE66D) & self._M

    def _next(self, bits: int) -> int:
        self.s = (self.s * 0x5DEECE66D + 0xB) & self._M
        return self.s >> (48 - bits)

    def next_double(self) -> float:
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))

    def next_float(self) -> float:
        return self._next(24) / float(1 << 24)

    def doubles(self, count: int) -> np.ndarray:
        return np.array([self.next_double() for _ in range(count)], dtype=np.float64)


def fill_uniform(count: int, seed: int = 2, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """Counter-based uniform fill shared by host oracle and device fill kernels.

    u(i) = (splitmix64(seed + i) >> 11) * 2^-53  (SURVEY.md section 8(d)).
    """
    i = np.arange(count, dtype=np.uint64) + np.uint64(seed)
    with np.errstate(over="ignore"):
        z = i * np.uint64(0x9E3779B97F4A7C15)  # golden-gamma stride then mix
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return lo + (hi - lo) * u


# --------------------------------------------------------------------------
# 1-D FFT  (fft/DoubleFFT_1D.java)
# --------------------------------------------------------------------------


def complex_forward_1d(a: np.ndarray, n: int, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.complexForward (fft/DoubleFFT_1D.java:243-263).

    X[k] = sum_j x[j] exp(-2 pi i j k / n); interleaved, in place, unscaled.
    n == 1 is a no-op (:248-250).
    """
    a = np.array(a, dtype=np.float64, copy=True)
    seg = a[offa:offa + 2 * n]
    a[offa:offa + 2 * n] = c2i(np.fft.fft(i2c(seg)))
    return a


def complex_inverse_1d(a: np.ndarray, n: int, scale: bool, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.complexInverse (fft/DoubleFFT_1D.java:362-385); scale => 1/n."""
    a = np.array(a, dtype=np.float64, copy=True)
    z = np.fft.ifft(i2c(a[offa:offa + 2 * n]))
    if not scale:
        z = z * n
    a[offa:offa + 2 * n] = c2i(z)
    return a


def pack_real_1d(X: np.ndarray, n: int) -> np.ndarray:
    """Packed half-spectrum layout of realForward (fft/DoubleFFT_1D.java:436-450).

    even n: a[2k]=Re[k] (0<=k<n/2), a[2k+1]=Im[k] (0<k<n/2), a[1]=Re[n/2];
    odd  n: a[2k]=Re[k] (0<=k<(n+1)/2), a[2k+1]=Im[k] (0<k<(n-1)/2),
            a[1]=Im[(n-1)/2].
    """
    out = np.zeros(n, dtype=np.float64)
    if n == 1:
        out[0] = X[0].real
        return out
    if n % 2 == 0:
        h = n // 2
        out[0:n:2] = X[0:h].real
        out[3:n:2] = X[1:h].imag
        out[1] = X[h].real
    else:
        h = (n + 1) // 2
        out[0:n:2] = X[0:h].real
        k = np.arange(1, (n - 1) // 2)
        out[2 * k + 1] = X[k].imag
        out[1] = X[(n - 1) // 2].imag
    return out


def unpack_real_1d(p: np.ndarray, n: int) -> np.ndarray:
    """Inverse of pack_real_1d: returns the half spectrum X[0..n/2] (numpy rfft shape)."""
    p = np.asarray(p, dtype=np.float64)
    m = n // 2 + 1
    X = np.zeros(m, dtype=np.complex128)
    if n == 1:
        X[0] = p[0]
        return X
    if n % 2 == 0:
        h = n // 2
        X[0:h] = p[0:n:2]
        X[1:h] += 1j * p[3:n:2]
        X[h] = p[1]
    else:
        h = (n + 1) // 2
        X[0:h] = p[0:n:2]
        k = np.arange(1, (n - 1) // 2)
        X[k] += 1j * p[2 * k + 1]
        X[(n - 1) // 2] += 1j * p[1]
    return X


def real_forward_1d(a: np.ndarray, n: int, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.realForward (fft/DoubleFFT_1D.java:524-561)."""
    a = np.array(a, dtype=np.float64, copy=True)
    if n == 1:
        return a
    X = np.fft.fft(a[offa:offa + n])
    a[offa:offa + n] = pack_real_1d(X, n)
    return a


def real_forward_full_1d(a: np.ndarray, n: int, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.realForwardFull (fft/DoubleFFT_1D.java:678-755).

    Input: n reals in a[offa:offa+n]; output: full interleaved spectrum in
    a[offa:offa+2n].  (The reference's pow2 branch leaves a[offa+n+1] untouched,
    :716-726, assuming it was 0; the true value Im[n/2] is 0 and we write it.)
    """
    a = np.array(a, dtype=np.float64, copy=True)
    X = np.fft.fft(a[offa:offa + n])
    a[offa:offa + 2 * n] = c2i(X)
    return a


def real_inverse_1d(a: np.ndarray, n: int, scale: bool, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.realInverse (fft/DoubleFFT_1D.java:946-989).

    Input in the packed layout.  Unscaled result is (n/2)*x for power-of-two n
    (scale factor 1/(n/2), :965) but n*x for the other plans (:977,:983) --
    pinned by src/test/java/org/jtransforms/fft/DoubleFFT_1DTest.java:573-596.
    """
    a = np.array(a, dtype=np.float64, copy=True)
    if n == 1:
        return a
    X = unpack_real_1d(a[offa:offa + n], n)
    x = np.fft.irfft(X, n)  # = true inverse (already 1/n)
    if not scale:
        x = x * (n / 2.0 if is_pow2(n) else float(n))
    a[offa:offa + n] = x
    return a


def real_inverse_full_1d(a: np.ndarray, n: int, scale: bool, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.realInverseFull (fft/DoubleFFT_1D.java:1112-1195):
    complexInverse of the real data with zero imaginary part, full 2n output."""
    a = np.array(a, dtype=np.float64, copy=True)
    z = np.fft.ifft(a[offa:offa + n])
    if not scale:
        z = z * n
    a[offa:offa + 2 * n] = c2i(z)
    return a


def real_inverse2_1d(a: np.ndarray, n: int, scale: bool, offa: int = 0) -> np.ndarray:
    """DoubleFFT_1D.realInverse2 (protected; fft/DoubleFFT_1D.java:1298-1357):
    packed half spectrum of the INVERSE DFT of real data (= realForward with
    negated imaginary parts), times 1/n when scale."""
    a = np.array(a, dtype=np.float64, copy=True)
    if n == 1:
        return a
    X = np.conj(np.fft.fft(a[offa:offa + n]))
    if scale:
        X = X / n
    a[offa:offa + n] = pack_real_1d(X, n)
    return a


# --------------------------------------------------------------------------
# 2-D FFT  (fft/DoubleFFT_2D.java)
# --------------------------------------------------------------------------


def complex_forward_2d(a, rows, cols):
    """DoubleFFT_2D.complexForward (fft/DoubleFFT_2D.java:115-213); a[r*2C+2c]."""
    z = i2c(np.asarray(a, dtype=np.float64).reshape(rows, 2 * cols))
    return c2i(np.fft.fft2(z)).ravel()


def complex_inverse_2d(a, rows, cols, scale):
    """DoubleFFT_2D.complexInverse (fft/DoubleFFT_2D.java:456-553)."""
    z = np.fft.ifft2(i2c(np.asarray(a, dtype=np.float64).reshape(rows, 2 * cols)))
    if not scale:
        z = z * (rows * cols)
    return c2i(z).ravel()


def pack_real_2d(F: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """Packed layout of DoubleFFT_2D.realForward (doc fft/DoubleFFT_2D.java:794-810).

    F is the full rows x cols complex spectrum.
    """
    R, C = rows, cols
    a = np.zeros((R, C), dtype=np.float64)
    k2 = np.arange(1, C // 2)
    # 0<=k1<R, 0<k2<C/2
    a[:, 2 * k2] = F[:, k2].real
    a[:, 2 * k2 + 1] = F[:, k2].imag
    # 0<k1<R/2 : column 0 and Nyquist column
    for k1 in range(1, R // 2):
        a[k1, 0] = F[k1, 0].real
        a[k1, 1] = F[k1, 0].imag
        a[R - k1, 1] = F[k1, C // 2].real
        a[R - k1, 0] = -F[k1, C // 2].imag
    a[0, 0] = F[0, 0].real
    a[0, 1] = F[0, C // 2].real
    a[R // 2, 0] = F[R // 2, 0].real
    a[R // 2, 1] = F[R // 2, C // 2].real
    return a.ravel()


def real2d_get_index(rows: int, cols: int, r: int, c: int):
    """RealFFTUtils_2D.getIndex (fft/RealFFTUtils_2D.java:191-242), restated.

    (r, c) addresses the *interleaved* full spectrum (c in [0, 2*cols)); returns
    index >= 0 (value), negative (negated value at -index) or None (zero).
    """
    cmod2 = c & 1
    rmul2 = r << 1
    if r != 0:
        if c <= 1:
            if rmul2 == rows:
                return None if cmod2 == 1 else (rows * cols) >> 1
            if rmul2 < rows:
                return cols * r + cmod2
            if cmod2 == 0:
                return cols * (rows - r)
            return -(cols * (rows - r) + 1)
        if c == cols or c == cols + 1:
            if rmul2 == rows:
                return None if cmod2 == 1 else ((rows * cols) >> 1) + 1
            if rmul2 < rows:
                if cmod2 == 0:
                    return cols * (rows - r) + 1
                return -(cols * (rows - r))
            return cols * r + 1 - cmod2
        if c < cols:
            return cols * r + c
        if cmod2 == 0:
            return cols * (rows + 2 - r) - c
        return -(cols * (rows + 2 - r) - c + 2)
    if c == 1 or c == cols + 1:
        return None
    if c == cols:
        return 1
    if c < cols:
        return c
    if cmod2 == 0:
        return (cols << 1) - c
    return -((cols << 1) - c + 2)


def unpack_real_2d(p: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """RealFFTUtils_2D.unpack over all (r, c): packed -> full interleaved rows x 2cols."""
    p = np.asarray(p, dtype=np.float64).ravel()
    out = np.zeros((rows, 2 * cols), dtype=np.float64)
    for r in range(rows):
        for c in range(2 * cols):
            i = real2d_get_index(rows, cols, r, c)
            if i is None:
                out[r, c] = 0.0
            elif i >= 0:
                out[r, c] = p[i]
            else:
                out[r, c] = -p[-i]
    return out


def real_forward_2d(a, rows, cols):
    """DoubleFFT_2D.realForward (fft/DoubleFFT_2D.java:820-838); pow2 only (:822-823)."""
    if not (is_pow2(rows) and is_pow2(cols)):
        raise ValueError("rows and columns must be power of two numbers")
    x = np.asarray(a, dtype=np.float64).reshape(rows, cols)
    return pack_real_2d(np.fft.fft2(x), rows, cols)


def sim_real_forward_2d(a, rows, cols):
    """Step-by-step restatement of the reference's pow2 algorithm, used only to
    pin pack_real_2d: rows realForward (fft/DoubleFFT_2D.java:830-832), complex
    column FFTs over all cols/2 interleaved columns including pseudo column 0
    (cdft2d_sub :2635-2792), then rdft2d_sub(1) (:2563-2573)."""
    x = np.array(a, dtype=np.float64).reshape(rows, cols)
    for r in range(rows):
        x[r] = real_forward_1d(x[r], cols)
    z = i2c(x)
    z = np.fft.fft(z, axis=0)
    x = c2i(z)
    for i in range(1, rows // 2):
        j = rows - i
        x[j, 0] = 0.5 * (x[i, 0] - x[j, 0])
        x[i, 0] -= x[j, 0]
        x[j, 1] = 0.5 * (x[i, 1] + x[j, 1])
        x[i, 1] -= x[j, 1]
    return x.ravel()


def real_forward_full_2d(a, rows, cols):
    """DoubleFFT_2D.realForwardFull (fft/DoubleFFT_2D.java:956-975, fillSymmetric
    :3877-3992, mixedRadixRealForwardFull :1509-1716): input rows*cols reals in
    the first half of a (row length cols); output rows x 2cols interleaved."""
    x = np.asarray(a, dtype=np.float64).ravel()[: rows * cols].reshape(rows, cols)
    return c2i(np.fft.fft2(x)).ravel()


def real_inverse_2d(a, rows, cols, scale):
    """DoubleFFT_2D.realInverse (fft/DoubleFFT_2D.java:1077-1095): packed input.

    Unscaled result is (rows*cols/2)*x: the column pass is a complexInverse
    (factor rows) and the row pass the pow2 1-D realInverse (factor cols/2).
    """
    if not (is_pow2(rows) and is_pow2(cols)):
        raise ValueError("rows and columns must be power of two numbers")
    full = i2c(unpack_real_2d(a, rows, cols))
    x = np.fft.ifft2(full).real
    if not scale:
        x = x * (rows * cols / 2.0)
    return x.ravel()


def real_inverse_full_2d(a, rows, cols, scale):
    """DoubleFFT_2D.realInverseFull (fft/DoubleFFT_2D.java:1212-1289)."""
    x = np.asarray(a, dtype=np.float64).ravel()[: rows * cols].reshape(rows, cols)
    z = np.fft.ifft2(x)
    if not scale:
        z = z * (rows * cols)
    return c2i(z).ravel()


# --------------------------------------------------------------------------
# 3-D FFT  (fft/DoubleFFT_3D.java)
# --------------------------------------------------------------------------


def complex_forward_3d(a, slices, rows, cols):
    """DoubleFFT_3D.complexForward (fft/DoubleFFT_3D.java:145-325);
    a[k1*sliceStride + k2*rowStride + 2*k3] (doc :127-141)."""
    z = i2c(np.asarray(a, dtype=np.float64).reshape(slices, rows, 2 * cols))
    return c2i(sfft.fftn(z, workers=_workers(z))).ravel()


def complex_inverse_3d(a, slices, rows, cols, scale):
    """DoubleFFT_3D.complexInverse (fft/DoubleFFT_3D.java:742-935)."""
    z = np.fft.ifftn(i2c(np.asarray(a, dtype=np.float64).reshape(slices, rows, 2 * cols)))
    if not scale:
        z = z * (slices * rows * cols)
    return c2i(z).ravel()


def pack_real_3d(F: np.ndarray, slices: int, rows: int, cols: int) -> np.ndarray:
    """Packed layout of DoubleFFT_3D.realForward (doc fft/DoubleFFT_3D.java:1298-1328)."""
    S, R, C = slices, rows, cols
    a = np.zeros((S, R, C), dtype=np.float64)
    k3 = np.arange(1, C // 2)
    a[:, :, 2 * k3] = F[:, :, k3].real
    a[:, :, 2 * k3 + 1] = F[:, :, k3].imag
    h = C // 2
    for k1 in range(S):
        m1 = (S - k1) % S
        for k2 in range(1, R // 2):
            a[k1, k2, 0] = F[k1, k2, 0].real
            a[k1, k2, 1] = F[k1, k2, 0].imag
            a[k1, R - k2, 1] = F[m1, k2, h].real
            a[k1, R - k2, 0] = -F[m1, k2, h].imag
    for k1 in range(1, S // 2):
        a[k1, 0, 0] = F[k1, 0, 0].real
        a[k1, 0, 1] = F[k1, 0, 0].imag
        a[k1, R // 2, 0] = F[k1, R // 2, 0].real
        a[k1, R // 2, 1] = F[k1, R // 2, 0].imag
        a[S - k1, 0, 1] = F[k1, 0, h].real
        a[S - k1, 0, 0] = -F[k1, 0, h].imag
        a[S - k1, R // 2, 1] = F[k1, R // 2, h].real
        a[S - k1, R // 2, 0] = -F[k1, R // 2, h].imag
    for k1 in (0, S // 2):
        for k2 in (0, R // 2):
            a[k1, k2, 0] = F[k1, k2, 0].real
            a[k1, k2, 1] = F[k1, k2, h].real
    return a.ravel()


def sim_real_forward_3d(a, slices, rows, cols):
    """Step-by-step restatement of the reference's pow2 3-D real algorithm
    (xdft3da_sub1 + cdft3db_sub + rdft3d_sub(1), fft/DoubleFFT_3D.java:1339-1355,
    :6909-7021), used only to pin pack_real_3d."""
    S, R, C = slices, rows, cols
    x = np.array(a, dtype=np.float64).reshape(S, R, C)
    for s in range(S):
        for r in range(R):
            x[s, r] = real_forward_1d(x[s, r], C)
    z = i2c(x)
    z = np.fft.fft(z, axis=1)
    z = np.fft.fft(z, axis=0)
    x = c2i(z)
    n1h, n2h = S >> 1, R >> 1

    def fix(p, q):  # a[p] = .5*(a[q]-a[p]) etc. on the (re, im) pair at col 0/1
        x[p][0] = 0.5 * (x[q][0] - x[p][0])
        x[q][0] -= x[p][0]
        x[p][1] = 0.5 * (x[q][1] + x[p][1])
        x[q][1] -= x[p][1]

    for i in range(1, n1h):
        j = S - i
        fix((j, 0), (i, 0))
        fix((j, n2h), (i, n2h))
        for k in range(1, n2h):
            l = R - k
            fix((j, l), (i, k))
            fix((i, l), (j, k))
    for k in range(1, n2h):
        l = R - k
        fix((0, l), (0, k))
        fix((n1h, l), (n1h, k))
    return x.ravel()


def unpack_real_3d(p, slices, rows, cols) -> np.ndarray:
    """Packed 3-D -> full complex spectrum (S, R, C) using Hermitian symmetry
    (inverse of pack_real_3d; the role RealFFTUtils_3D.unpack plays,
    fft/RealFFTUtils_3D.java:232-330)."""
    S, R, C = slices, rows, cols
    a = np.asarray(p, dtype=np.float64).reshape(S, R, C)
    h = C // 2
    F = np.zeros((S, R, C), dtype=np.complex128)
    k3 = np.arange(1, h)
    F[:, :, k3] = a[:, :, 2 * k3] + 1j * a[:, :, 2 * k3 + 1]
    have0 = np.zeros((S, R), dtype=bool)
    haveh = np.zeros((S, R), dtype=bool)
    for k1 in range(S):
        m1 = (S - k1) % S
        for k2 in range(1, R // 2):
            F[k1, k2, 0] = a[k1, k2, 0] + 1j * a[k1, k2, 1]
            have0[k1, k2] = True
            F[m1, k2, h] = a[k1, R - k2, 1] - 1j * a[k1, R - k2, 0]
            haveh[m1, k2] = True
    for k1 in range(1, S // 2):
        for k2 in (0, R // 2):
            F[k1, k2, 0] = a[k1, k2, 0] + 1j * a[k1, k2, 1]
            have0[k1, k2] = True
            F[k1, k2, h] = a[S - k1, k2, 1] - 1j * a[S - k1, k2, 0]
            haveh[k1, k2] = True
    for k1 in (0, S // 2):
        for k2 in (0, R // 2):
            F[k1, k2, 0] = a[k1, k2, 0]
            F[k1, k2, h] = a[k1, k2, 1]
            have0[k1, k2] = haveh[k1, k2] = True
    for k1 in range(S):
        for k2 in range(R):
            m1, m2 = (S - k1) % S, (R - k2) % R
            if not have0[k1, k2]:
                F[k1, k2, 0] = np.conj(F[m1, m2, 0])
            if not haveh[k1, k2]:
                F[k1, k2, h] = np.conj(F[m1, m2, h])
    # upper half of k3 by symmetry
    for k in range(h + 1, C):
        F[:, :, k] = np.conj(np.roll(np.roll(F[::-1, ::-1, C - k], 1, axis=0), 1, axis=1))
    return F


def real_forward_3d(a, slices, rows, cols):
    """DoubleFFT_3D.realForward (fft/DoubleFFT_3D.java:1339-1355); pow2 only."""
    if not (is_pow2(slices) and is_pow2(rows) and is_pow2(cols)):
        raise ValueError("slices, rows and columns must be power of two numbers")
    x = np.asarray(a, dtype=np.float64).reshape(slices, rows, cols)
    return pack_real_3d(np.fft.fftn(x), slices, rows, cols)


def real_forward_full_3d(a, slices, rows, cols):
    """DoubleFFT_3D.realForwardFull (fft/DoubleFFT_3D.java:1480-1500,
    fillSymmetric :7387-7614)."""
    n = slices * rows * cols
    x = np.asarray(a, dtype=np.float64).ravel()[:n].reshape(slices, rows, cols)
    return c2i(np.fft.fftn(x)).ravel()


def real_inverse_3d(a, slices, rows, cols, scale):
    """DoubleFFT_3D.realInverse (fft/DoubleFFT_3D.java:1629-1645).  Unscaled
    factor is slices*rows*cols/2 (complex passes x pow2 1-D realInverse)."""
    if not (is_pow2(slices) and is_pow2(rows) and is_pow2(cols)):
        raise ValueError("slices, rows and columns must be power of two numbers")
    F = unpack_real_3d(a, slices, rows, cols)
    x = np.fft.ifftn(F).real
    if not scale:
        x = x * (slices * rows * cols / 2.0)
    return x.ravel()


def real_inverse_full_3d(a, slices, rows, cols, scale):
    """DoubleFFT_3D.realInverseFull (fft/DoubleFFT_3D.java:1770-1790)."""
    n = slices * rows * cols
    x = np.asarray(a, dtype=np.float64).ravel()[:n].reshape(slices, rows, cols)
    z = np.fft.ifftn(x)
    if not scale:
        z = z * n
    return c2i(z).ravel()


# --------------------------------------------------------------------------
# DCT / DST / DHT  (dct/, dst/, dht/)
# --------------------------------------------------------------------------


def dct_forward_1d(x, scale: bool):
    """DoubleDCT_1D.forward (dct/DoubleDCT_1D.java:169-243), DCT-II along the
    last axis.

    pow2 (Ooura ddct, :176-194): unscaled C[k] = sum_j x[j] cos(pi (j+1/2) k/n);
    non-pow2 (mirror to 2n + realForward + twiddle, :196-241): unscaled value is
    2x that.  scale=True is orthonormal in both branches (:190-193, :237-240).
    n == 1 is a no-op (:171-173).
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    if n == 1:
        return x.copy()
    if scale:
        return sfft.dct(x, type=2, norm="ortho", axis=-1, workers=_workers(x))
    y = sfft.dct(x, type=2, axis=-1, workers=_workers(x))  # scipy: 2*sum
    return y * 0.5 if is_pow2(n) else y


def dct_inverse_1d(x, scale: bool):
    """DoubleDCT_1D.inverse (dct/DoubleDCT_1D.java:361-434), DCT-III, last axis.

    pow2 unscaled (:368-387): y[j] = sum_k a[k] cos(pi k (j+1/2)/n), k=0 at full
    weight.  non-pow2 unscaled (:389-431): y[j] = (1/n)(a[0]/2 + sum_{k>=1} ...).
    scale=True is the orthonormal inverse in both branches.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    if n == 1:
        return x.copy()
    if scale:
        return sfft.idct(x, type=2, norm="ortho", axis=-1)
    # scipy dct type 3 unnormalised: y[j] = a0 + 2 sum_{k>=1} a_k cos(pi k (2j+1)/(2n))
    y3 = sfft.dct(x, type=3, axis=-1)
    a0 = x[..., 0:1]
    if is_pow2(n):
        return 0.5 * (y3 + a0)            # a0 + sum_{k>=1}
    return y3 / (2.0 * n)                  # (a0/2 + sum_{k>=1}) / n


def _dst_pre(x):
    y = np.array(x, dtype=np.float64, copy=True)
    y[..., 1::2] = -y[..., 1::2]
    return y


def dst_forward_1d(x, scale: bool):
    """DoubleDST_1D.forward (dst/DoubleDST_1D.java:96-160): negate odd inputs,
    DCT-II, reverse.  => S[k] = sum_j x[j] sin(pi (j+1/2)(k+1)/n) (pow2, unscaled)."""
    x = np.asarray(x, dtype=np.float64)
    if x.shape[-1] == 1:
        return x.copy()
    return dct_forward_1d(_dst_pre(x), scale)[..., ::-1].copy()


def dst_inverse_1d(x, scale: bool):
    """DoubleDST_1D.inverse (dst/DoubleDST_1D.java:264-325): reverse, DCT-III,
    negate odd outputs."""
    x = np.asarray(x, dtype=np.float64)
    if x.shape[-1] == 1:
        return x.copy()
    return _dst_pre(dct_inverse_1d(x[..., ::-1], scale))


def dht_forward_1d(x):
    """DoubleDHT_1D.forward (dht/DoubleDHT_1D.java:94-152):
    H[k] = sum_j x[j] cas(2 pi j k/n) = Re X[k] - Im X[k]."""
    x = np.asarray(x, dtype=np.float64)
    if x.shape[-1] == 1:
        return x.copy()
    X = np.fft.fft(x, axis=-1)
    return X.real - X.imag


def dht_inverse_1d(x, scale: bool):
    """DoubleDHT_1D.inverse (dht/DoubleDHT_1D.java:255-270): forward, then 1/n if scale."""
    y = dht_forward_1d(x)
    n = np.asarray(x).shape[-1]
    return y / n if (scale and n > 1) else y


def _apply_axes(fn, x, axes):
    for ax in axes:
        x = np.moveaxis(fn(np.moveaxis(x, ax, -1)), -1, ax)
    return x


def dct_forward_nd(a, shape, scale):
    """DoubleDCT_2D/3D.forward (dct/DoubleDCT_2D.java:104-183,
    dct/DoubleDCT_3D.java:133-…): separable, each axis with the 1-D rule of its
    own length."""
    x = np.asarray(a, dtype=np.float64).reshape(shape)
    return _apply_axes(lambda v: dct_forward_1d(v, scale), x, range(len(shape))).ravel()


def dct_inverse_nd(a, shape, scale):
    """DoubleDCT_2D/3D.inverse (dct/DoubleDCT_2D.java:360-440)."""
    x = np.asarray(a, dtype=np.float64).reshape(shape)
    return _apply_axes(lambda v: dct_inverse_1d(v, scale), x, range(len(shape))).ravel()


def dst_forward_nd(a, shape, scale):
    """DoubleDST_2D/3D.forward (dst/DoubleDST_2D.java:103-…)."""
    x = np.asarray(a, dtype=np.float64).reshape(shape)
    return _apply_axes(lambda v: dst_forward_1d(v, scale), x, range(len(shape))).ravel()


def dst_inverse_nd(a, shape, scale):
    """DoubleDST_2D/3D.inverse."""
    x = np.asarray(a, dtype=np.float64).reshape(shape)
    return _apply_axes(lambda v: dst_inverse_1d(v, scale), x, range(len(shape))).ravel()


def dht_forward_nd(a, shape):
    """DoubleDHT_2D/3D.forward (dht/DoubleDHT_2D.java:102-190 + yTransform
    :1288-1309; dht/DoubleDHT_3D.java yTransform :2314-2356): separable 1-D DHTs
    followed by yTransform, which yields the true multi-dimensional DHT
    H = Re(F) - Im(F) with F = fftn(x) (checked in tests/test_oracle.py against a
    literal restatement of yTransform, ``sim_dht_nd``)."""
    x = np.asarray(a, dtype=np.float64).reshape(shape)
    F = sfft.fftn(x, workers=_workers(x))
    return (F.real - F.imag).ravel()


def dht_inverse_nd(a, shape, scale):
    """DoubleDHT_2D/3D.inverse (dht/DoubleDHT_2D.java:360-…): forward then
    1/(product of sizes) when scale."""
    y = dht_forward_nd(a, shape)
    return y / float(np.prod(shape)) if scale else y


def sim_dht_nd(a, shape):
    """Literal restatement of separable DHT + yTransform (2-D:
    dht/DoubleDHT_2D.java:1288-1309, 3-D: dht/DoubleDHT_3D.java:2314-2356)."""
    x = np.array(a, dtype=np.float64).reshape(shape)
    x = _apply_axes(dht_forward_1d, x, range(len(shape)))
    if len(shape) == 2:
        R, C = shape
        for r in range(R // 2 + 1):
            mr = (R - r) % R
            for c in range(C // 2 + 1):
                mc = (C - c) % C
                A, B, Cc, D = x[r, c], x[mr, c], x[r, mc], x[mr, mc]
                E = ((A + D) - (B + Cc)) / 2
                x[r, c] = A - E
                x[mr, c] = B + E
                x[r, mc] = Cc + E
                x[mr, mc] = D - E
    else:
        S, R, C = shape
        for s in range(S // 2 + 1):
            sC = (S - s) % S
            for r in range(R // 2 + 1):
                rC = (R - r) % R
                for c in range(C // 2 + 1):
                    cC = (C - c) % C
                    i1, i2, i3, i4 = (s, rC, c), (s, r, cC), (sC, r, c), (sC, rC, cC)
                    i5, i6, i7, i8 = (sC, rC, c), (sC, r, cC), (s, r, c), (s, rC, cC)
                    A, B, Cv, D = x[i1], x[i2], x[i3], x[i4]
                    E, F, G, H = x[i5], x[i6], x[i7], x[i8]
                    x[i7] = (A + B + Cv - D) / 2
                    x[i3] = (E + F + G - H) / 2
                    x[i1] = (G + H + E - F) / 2
                    x[i5] = (Cv + D + A - B) / 2
                    x[i2] = (H + G + F - E) / 2
                    x[i6] = (D + Cv + B - A) / 2
                    x[i8] = (B + A + D - Cv) / 2
                    x[i4] = (F + E + H - G) / 2
    return x.ravel()


def bluestein_forward_sim(x: np.ndarray) -> np.ndarray:
    """Restatement of the reference's chirp-z data flow
    (bluesteini fft/DoubleFFT_1D.java:1864-1890, bluestein_complex :1920-2107):
    bk1[i] = exp(+i pi (i^2 mod 2n)/n); bk2 = FFT_M(wrap(bk1)/M);
    ak = x*conj(bk1) zero-padded to M = nextPow2(2n-1); FFT_M; *bk2; unscaled
    inverse FFT_M; *conj(bk1).  Used to pin that data flow == DFT."""
    x = np.asarray(x, dtype=np.complex128)
    n = len(x)
    M = next_pow2(2 * n - 1)
    i = np.arange(n, dtype=np.int64)
    ph = (i * i) % (2 * n)
    bk1 = np.exp(1j * np.pi * ph / n)
    wrap = np.zeros(M, dtype=np.complex128)
    wrap[:n] = bk1 / M
    wrap[M - n + 1:] = bk1[1:][::-1] / M
    bk2 = np.fft.fft(wrap)
    ak = np.zeros(M, dtype=np.complex128)
    ak[:n] = x * np.conj(bk1)
    ak = np.fft.fft(ak) * bk2
    ak = np.fft.ifft(ak) * M
    return ak[:n] * np.conj(bk1)
