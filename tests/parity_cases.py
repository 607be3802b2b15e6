"""Parity cases shared by the CPU (emulated-kernel) and GPU test modules.

Every case drives the public org.jtransforms-style API of ``jtransforms_b200`` and compares the result
with the oracle (oracle/jt_oracle.py) on the same seeded input.  Tolerances are the north-star ones:
relative L2 <= 1e-12*log2(N) for double and 1e-5*log2(N) for float (N = total points, at least 2).
The structure follows the reference's own tests (src/test/java/org/jtransforms/fft/DoubleFFT_1DTest.java
:208-596, DoubleFFT_2DTest.java:167-476, DoubleFFT_3DTest.java:165-199, dct/DoubleDCT_1DTest.java:125-160).
"""
from __future__ import annotations

import math

import numpy as np

from oracle import jt_oracle as o


def tol(prec: str, total: int) -> float:
    lg = max(1.0, math.log2(max(2, total)))
    return (1e-12 if prec == "Double" else 1e-5) * lg


def dtype_of(prec):
    return np.float64 if prec == "Double" else np.float32


def rnd(count, seed=20110602, lo=-1.0, hi=1.0):
    return o.fill_uniform(count, seed=seed, lo=lo, hi=hi)


def check(got, want, prec, total, what=""):
    err = o.rel_l2(np.asarray(got, dtype=np.float64), want)
    assert err <= tol(prec, total), "%s: rel L2 %.3e > %.3e" % (what, err, tol(prec, total))
    return err


def cls(jt, prec, name):
    return getattr(jt, prec + name)


# ------------------------------------------------------------------ 1-D FFT
def fft1d_complex(jt, prec, n, offa=0):
    dt = dtype_of(prec)
    x = rnd(2 * n + offa).astype(dt)
    f = cls(jt, prec, "FFT_1D")(n)
    a = x.copy()
    f.complexForward(a, offa) if offa else f.complexForward(a)
    want = o.complex_forward_1d(x.astype(np.float64), n, offa)
    check(a, want, prec, n, "complexForward n=%d" % n)
    for scale in (True, False):
        b = x.copy()
        f.complexInverse(b, offa, scale)
        check(b, o.complex_inverse_1d(x.astype(np.float64), n, scale, offa), prec, n, "complexInverse n=%d" % n)


def fft1d_real(jt, prec, n):
    dt = dtype_of(prec)
    x = rnd(n).astype(dt)
    f = cls(jt, prec, "FFT_1D")(n)
    a = x.copy()
    f.realForward(a)
    want = o.real_forward_1d(x.astype(np.float64), n)
    check(a, want, prec, n, "realForward n=%d" % n)
    for scale in (True, False):
        b = want.astype(dt)
        f.realInverse(b, scale)
        check(b, o.real_inverse_1d(want, n, scale), prec, n, "realInverse n=%d scale=%s" % (n, scale))
    a2 = np.zeros(2 * n, dtype=dt)
    a2[:n] = x
    f.realForwardFull(a2)
    z = np.zeros(2 * n)
    z[:n] = x
    check(a2, o.real_forward_full_1d(z, n), prec, n, "realForwardFull n=%d" % n)
    for scale in (True, False):
        a3 = np.zeros(2 * n, dtype=dt)
        a3[:n] = x
        f.realInverseFull(a3, scale)
        check(a3, o.real_inverse_full_1d(z, n, scale), prec, n, "realInverseFull n=%d" % n)


def fft1d_batch(jt, prec, n, howmany, pad=0):
    dt = dtype_of(prec)
    dist = 2 * n + pad
    x = rnd(howmany * dist).astype(dt)
    a = x.copy()
    cls(jt, prec, "FFT_1D")(n).complexForwardBatch(a, howmany, dist)
    want = x.astype(np.float64).copy()
    for b in range(howmany):
        want[b * dist:b * dist + 2 * n] = o.complex_forward_1d(want[b * dist:b * dist + 2 * n], n)
    check(a, want, prec, n, "batch n=%d x%d" % (n, howmany))


# ------------------------------------------------------------------ 2-D / 3-D FFT
def fftnd_complex(jt, prec, dims):
    dt = dtype_of(prec)
    total = int(np.prod(dims))
    x = rnd(2 * total, lo=0.0, hi=1.0).astype(dt)
    f = cls(jt, prec, "FFT_%dD" % len(dims))(*dims)
    a = x.copy()
    f.complexForward(a)
    fwd = o.complex_forward_2d if len(dims) == 2 else o.complex_forward_3d
    inv = o.complex_inverse_2d if len(dims) == 2 else o.complex_inverse_3d
    check(a, fwd(x.astype(np.float64), *dims), prec, total, "complexForward %s" % (dims,))
    for scale in (True, False):
        b = x.copy()
        f.complexInverse(b, scale)
        check(b, inv(x.astype(np.float64), *dims, scale), prec, total, "complexInverse %s" % (dims,))


def fftnd_real(jt, prec, dims):
    dt = dtype_of(prec)
    total = int(np.prod(dims))
    x = rnd(total, lo=0.0, hi=1.0).astype(dt)
    f = cls(jt, prec, "FFT_%dD" % len(dims))(*dims)
    two = len(dims) == 2
    a = x.copy()
    f.realForward(a)
    want = (o.real_forward_2d if two else o.real_forward_3d)(x.astype(np.float64), *dims)
    check(a, want, prec, total, "realForward %s" % (dims,))
    for scale in (True, False):
        b = want.astype(dt)
        f.realInverse(b, scale)
        check(b, (o.real_inverse_2d if two else o.real_inverse_3d)(want, *dims, scale), prec, total,
              "realInverse %s" % (dims,))


def fftnd_real_full(jt, prec, dims):
    dt = dtype_of(prec)
    total = int(np.prod(dims))
    x = rnd(total, lo=0.0, hi=1.0).astype(dt)
    f = cls(jt, prec, "FFT_%dD" % len(dims))(*dims)
    two = len(dims) == 2
    a = np.zeros(2 * total, dtype=dt)
    a[:total] = x
    f.realForwardFull(a)
    check(a, (o.real_forward_full_2d if two else o.real_forward_full_3d)(x.astype(np.float64), *dims), prec, total,
          "realForwardFull %s" % (dims,))
    a = np.zeros(2 * total, dtype=dt)
    a[:total] = x
    f.realInverseFull(a, True)
    check(a, (o.real_inverse_full_2d if two else o.real_inverse_full_3d)(x.astype(np.float64), *dims, True), prec,
          total, "realInverseFull %s" % (dims,))


# ------------------------------------------------------------------ DCT / DST / DHT
def r2r(jt, prec, kind, dims):
    dt = dtype_of(prec)
    total = int(np.prod(dims))
    x = rnd(total).astype(dt)
    x64 = x.astype(np.float64)
    t = cls(jt, prec, "%s_%dD" % (kind, len(dims)))(*dims)
    shape = tuple(dims)
    for scale in (True, False):
        a = x.copy()
        if kind == "DHT":
            if not scale:
                continue
            t.forward(a)
            want = o.dht_forward_nd(x64, shape)
        else:
            t.forward(a, scale)
            want = (o.dct_forward_nd if kind == "DCT" else o.dst_forward_nd)(x64, shape, scale)
        check(a, want, prec, total, "%s forward %s scale=%s" % (kind, dims, scale))
    for scale in (True, False):
        a = x.copy()
        t.inverse(a, scale)
        if kind == "DHT":
            want = o.dht_inverse_nd(x64, shape, scale)
        else:
            want = (o.dct_inverse_nd if kind == "DCT" else o.dst_inverse_nd)(x64, shape, scale)
        check(a, want, prec, total, "%s inverse %s scale=%s" % (kind, dims, scale))
    # the reference's own (only) DCT/DST/DHT test: inverse(forward(x, true), true) == x
    a = x.copy()
    if kind == "DHT":
        t.forward(a)
    else:
        t.forward(a, True)
    t.inverse(a, True)
    check(a, x64, prec, total, "%s round trip %s" % (kind, dims))


# ------------------------------------------------------------------ multi-GPU plans through the host API
def fft3d_multi(jt, prec, dims, devices):
    """DoubleFFT_3D(..., devices=[...]): one host array in, natural order out (jtb_plan_set_devices + jtb_exec)"""
    S, R, Cn = dims
    total = S * R * Cn
    dt = dtype_of(prec)
    f = getattr(jt, prec + "FFT_3D")(S, R, Cn, devices=devices)
    x64 = rnd(2 * total)
    a = x64.astype(dt)
    f.complexForward(a)
    check(a, o.complex_forward_3d(x64, S, R, Cn), prec, total, "multi-GPU forward %s P=%d" % (dims, len(devices)))
    f.complexInverse(a, True)
    check(a, x64, prec, total, "multi-GPU round trip %s P=%d" % (dims, len(devices)))
    b = x64.astype(dt)
    f.complexInverse(b, False)
    check(b, o.complex_inverse_3d(x64, S, R, Cn, False), prec, total, "multi-GPU unscaled inverse %s" % (dims,))


# ------------------------------------------------------------------ fused k2 + exchange (virtual ranks)
def slab_scatter_virtual(lib, prec, dims, P, torch_device="cpu", dev_index=0, fused_slices=False):
    """Runs the slab-decomposed forward transform with P *virtual* ranks inside one process: every rank's
    k3 pass and fused k2-scatter (jtb_fft3d_k2_scatter) write into P receive buffers that live on the same
    device, then each rank's k1 pass runs on its buffer.  Checks the kernel's addressing against the oracle."""
    import ctypes as C
    import torch
    S, R, Cn = dims
    Ls, Rh = S // P, R // P
    dt = dtype_of(prec)
    tdt = torch.float64 if prec == "Double" else torch.float32
    pcode = 0 if prec == "Double" else 1
    x = rnd(2 * S * R * Cn, lo=0.0, hi=1.0).astype(dt)
    locs = [torch.from_numpy(x.reshape(S, -1)[g * Ls:(g + 1) * Ls].copy().ravel()).to(torch_device) for g in range(P)]
    recvs = [torch.zeros(2 * S * Rh * Cn, dtype=tdt, device=torch_device) for _ in range(P)]
    arr = (C.c_void_p * P)(*[r.data_ptr() for r in recvs])

    def st():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream) if torch_device != "cpu" else None
    for g in range(P):
        if fused_slices:
            # rows + columns + exchange through the single entry point (persistent fused kernel for 512^2 slices)
            rc = lib.jtb_fft2d_slices_device(pcode, dev_index, C.c_void_p(locs[g].data_ptr()), Ls, R, Cn, P, g, arr, 0,
                                             st())
            assert rc == 0, lib.jtb_last_error()
            continue
        assert lib.jtb_lines_c2c_device(pcode, dev_index, C.c_void_p(locs[g].data_ptr()), Cn, Ls * R, 1, 0, Cn, 1, 0,
                                        1.0, st()) == 0
        rc = lib.jtb_fft3d_k2_scatter(pcode, dev_index, C.c_void_p(locs[g].data_ptr()), Ls, R, Cn, P, g, arr, 0, st())
        assert rc == 0, lib.jtb_last_error()
    want = o.complex_forward_3d(x.astype(np.float64), S, R, Cn).reshape(S, R, 2 * Cn)
    for h in range(P):
        assert lib.jtb_lines_c2c_device(pcode, dev_index, C.c_void_p(recvs[h].data_ptr()), S, Rh * Cn, Rh * Cn, 1,
                                        S * Rh * Cn, Rh * Cn, 0, 1.0, st()) == 0
        got = recvs[h].cpu().numpy().reshape(S, Rh, 2 * Cn)
        check(got, want[:, h * Rh:(h + 1) * Rh], prec, S * R * Cn, "slab scatter P=%d rank %d" % (P, h))
    # the way back: fused k1 pass + exchange (jtb_fft3d_k1_scatter) into P slab buffers, then the local k2 / k3 passes
    backs = [torch.zeros(2 * Ls * R * Cn, dtype=tdt, device=torch_device) for _ in range(P)]
    barr = (C.c_void_p * P)(*[r.data_ptr() for r in backs])
    for h in range(P):
        rc = lib.jtb_fft3d_k1_scatter(pcode, dev_index, C.c_void_p(recvs[h].data_ptr()), S, Rh, Cn, P, h, barr, 1, st())
        if rc == 2:          # no fused kernel for this slice count
            assert S not in (64, 512), lib.jtb_last_error()
            return
        assert rc == 0, lib.jtb_last_error()
    sc = 1.0 / (S * R * Cn)
    for g in range(P):
        assert lib.jtb_lines_c2c_device(pcode, dev_index, C.c_void_p(backs[g].data_ptr()), R, Cn * Ls, Cn, 1, R * Cn, Cn,
                                        1, 1.0, st()) == 0
        assert lib.jtb_lines_c2c_device(pcode, dev_index, C.c_void_p(backs[g].data_ptr()), Cn, Ls * R, 1, 0, Cn, 1, 1, sc,
                                        st()) == 0
        # recvs hold the forward spectrum; the three inverse passes with the 1/(S R C) scale give the input back
        got = backs[g].cpu().numpy()
        ref = x.astype(np.float64).reshape(S, -1)[g * Ls:(g + 1) * Ls].ravel()
        check(got, ref, prec, S * R * Cn, "slab round trip P=%d rank %d" % (P, g))
