"""org.jtransforms.fft.RealFFTUtils_2D / RealFFTUtils_3D mirror: where a logical Fourier mode lives inside the
packed array written by ``realForward`` (fft/RealFFTUtils_2D.java:191-242, fft/RealFFTUtils_3D.java:232-330).

Pure index arithmetic on the host (no transform code), written from the layout rules of
fft/DoubleFFT_2D.java:794-810 and fft/DoubleFFT_3D.java:1298-1328 rather than as a branch ladder:
a mode is either stored directly, stored for its Hermitian mirror (conjugate), or structurally real/zero.

``getIndex`` keeps the reference's convention: ``index >= 0`` -> ``packed[index]``, ``index < 0`` ->
``-packed[-index]``, ``MIN_VALUE`` -> the entry is zero.  Column index ``c`` runs over the interleaved full
spectrum (``c = 2*k + 0/1`` for real/imaginary part).
"""
from __future__ import annotations

MIN_VALUE = -(1 << 63)


def _signed(idx, sign):
    if idx is None:
        return MIN_VALUE
    return idx if sign > 0 else -idx


class RealFFTUtils_2D:
    def __init__(self, rows: int, columns: int):
        self.rows, self.columns = int(rows), int(columns)

    def _mode(self, k1: int, k2: int):
        """(index of Re, sign, index of Im, sign); index None = zero"""
        R, C = self.rows, self.columns
        h = C // 2
        if k2 > h:                                   # F[k1][k2] = conj F[-k1][-k2]
            ire, sre, iim, sim = self._mode((R - k1) % R, C - k2)
            return ire, sre, iim, -sim
        if 0 < k2 < h:
            return k1 * C + 2 * k2, 1, k1 * C + 2 * k2 + 1, 1
        if 2 * k1 > R:                               # columns 0 and C/2 are Hermitian along k1
            ire, sre, iim, sim = self._mode(R - k1, k2)
            return ire, sre, iim, -sim
        if k1 == 0 or 2 * k1 == R:                   # purely real corners
            return k1 * C + (0 if k2 == 0 else 1), 1, None, 1
        if k2 == 0:
            return k1 * C, 1, k1 * C + 1, 1
        return (R - k1) * C + 1, 1, (R - k1) * C, -1  # Re[k1][C/2], -Im[k1][C/2] live in row R-k1

    def getIndex(self, r: int, c: int) -> int:
        ire, sre, iim, sim = self._mode(r, c >> 1)
        return _signed(ire, sre) if (c & 1) == 0 else _signed(iim, sim)

    def unpack(self, r: int, c: int, packed, pos: int = 0):
        i = self.getIndex(r, c)
        if i == MIN_VALUE:
            return 0.0
        return packed[pos + i] if i >= 0 else -packed[pos - i]

    def pack(self, val, r: int, c: int, packed, pos: int = 0):
        i = self.getIndex(r, c)
        if i == MIN_VALUE:
            raise ValueError("[%d][%d] component cannot be modified (always zero)" % (r, c))
        if i >= 0:
            packed[pos + i] = val
        else:
            packed[pos - i] = -val


class RealFFTUtils_3D:
    def __init__(self, slices: int, rows: int, columns: int):
        self.slices, self.rows, self.columns = int(slices), int(rows), int(columns)

    def _at(self, k1, k2, k3):
        return (k1 * self.rows + k2) * self.columns + k3

    def _mode(self, k1: int, k2: int, k3: int):
        S, R, C = self.slices, self.rows, self.columns
        h = C // 2
        if k3 > h:
            ire, sre, iim, sim = self._mode((S - k1) % S, (R - k2) % R, C - k3)
            return ire, sre, iim, -sim
        if 0 < k3 < h:
            return self._at(k1, k2, 2 * k3), 1, self._at(k1, k2, 2 * k3 + 1), 1
        # planes k3 = 0 and k3 = C/2 are Hermitian in (k1, k2)
        if 2 * k2 > R:
            ire, sre, iim, sim = self._mode((S - k1) % S, R - k2, k3)
            return ire, sre, iim, -sim
        if 0 < 2 * k2 < R:
            if k3 == 0:
                return self._at(k1, k2, 0), 1, self._at(k1, k2, 1), 1
            j1 = (S - k1) % S                        # stored in slice -k1, row R-k2
            return self._at(j1, R - k2, 1), 1, self._at(j1, R - k2, 0), -1
        # k2 in {0, R/2}: Hermitian along k1 alone
        if 2 * k1 > S:
            ire, sre, iim, sim = self._mode(S - k1, k2, k3)
            return ire, sre, iim, -sim
        if k1 == 0 or 2 * k1 == S:
            return self._at(k1, k2, 0 if k3 == 0 else 1), 1, None, 1
        if k3 == 0:
            return self._at(k1, k2, 0), 1, self._at(k1, k2, 1), 1
        return self._at(S - k1, k2, 1), 1, self._at(S - k1, k2, 0), -1

    def getIndex(self, s: int, r: int, c: int) -> int:
        ire, sre, iim, sim = self._mode(s, r, c >> 1)
        return _signed(ire, sre) if (c & 1) == 0 else _signed(iim, sim)

    def unpack(self, s: int, r: int, c: int, packed, pos: int = 0):
        i = self.getIndex(s, r, c)
        if i == MIN_VALUE:
            return 0.0
        return packed[pos + i] if i >= 0 else -packed[pos - i]

    def pack(self, val, s: int, r: int, c: int, packed, pos: int = 0):
        i = self.getIndex(s, r, c)
        if i == MIN_VALUE:
            raise ValueError("[%d][%d][%d] component cannot be modified (always zero)" % (s, r, c))
        if i >= 0:
            packed[pos + i] = val
        else:
            packed[pos - i] = -val
