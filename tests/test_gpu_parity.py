"""GPU parity suite: the nvcc-built libjtb200.so, called through the C ABI (ctypes), against the oracle."""
import os

import numpy as np
import pytest

import parity_cases as pc
from oracle import jt_oracle as o

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FFTW = os.path.join(HERE, "golden", "fftw")


@pytest.fixture(scope="module")
def jt():
    import jtransforms_b200 as m
    from jtransforms_b200 import _lib
    _lib._lib = None
    lib = _lib.get()            # raises if libjtb200.so is missing: no fallback
    assert lib.jtb_device_count() >= 1
    lib.jtb_debug_set_limits(0, 0)
    return m


def _sizes():
    with open(os.path.join(FFTW, "sizes.txt")) as f:
        return [int(s) for s in f.read().split()]


@pytest.mark.parametrize("n", _sizes())
def test_fftw_golden(jt, n):
    """src/test/java/org/jtransforms/fft/DoubleFFT_1DTest.java:208-219 and FloatFFT_1DTest.java:203"""
    x = np.fromfile(os.path.join(FFTW, "fftw%d.in" % n), dtype="<f8")
    want = np.fromfile(os.path.join(FFTW, "fftw%d.out" % n), dtype="<f8")
    a = x.copy()
    jt.DoubleFFT_1D(n).complexForward(a)
    pc.check(a, want, "Double", n, "fftw %d" % n)
    assert o.rmse(a, want) <= 1e-12 * max(1.0, np.sqrt(n))
    b = x.astype(np.float32)
    jt.FloatFFT_1D(n).complexForward(b)
    pc.check(b, want, "Float", n, "fftw float %d" % n)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 16, 100, 120, 211, 310, 512, 1024, 4096, 8192, 16384, 65536, 10158,
                               65530, 1 << 20, 1 << 21])
def test_fft1d_complex(jt, prec, n):
    pc.fft1d_complex(jt, prec, n)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("n", [2, 4, 5, 9, 64, 101, 1024, 4096, 16384, 32768, 65536, 100000])
def test_fft1d_real(jt, prec, n):
    pc.fft1d_real(jt, prec, n)


def test_fft1d_bluestein_prime_batch(jt):
    """config 3 (reduced batch): FloatFFT_1D, n = 1 000 003 (prime), Bluestein"""
    pc.fft1d_batch(jt, "Float", 1000003, 3, pad=2)
    pc.fft1d_complex(jt, "Double", 1000003)


def test_fft1d_batch(jt):
    pc.fft1d_batch(jt, "Double", 4096, 37)
    pc.fft1d_batch(jt, "Float", 1 << 15, 5, pad=6)
    pc.fft1d_batch(jt, "Double", 1000, 9)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("dims", [(2, 2), (64, 128), (100, 120), (1024, 512), (4096, 64), (16, 8192), (311, 64), (2048, 128),
                                  (8192, 32)])
def test_fft2d_complex(jt, prec, dims):
    pc.fftnd_complex(jt, prec, dims)


@pytest.mark.parametrize("dims", [(2, 2), (2, 16), (16, 2), (256, 512), (1024, 64), (4096, 4096)])
def test_fft2d_real(jt, dims):
    """config 2 at full size is the last case"""
    pc.fftnd_real(jt, "Double", dims)


def test_fft2d_real_misc(jt):
    pc.fftnd_real(jt, "Float", (128, 256))
    pc.fftnd_real_full(jt, "Double", (64, 32))
    pc.fftnd_real_full(jt, "Double", (30, 50))


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("dims", [(2, 2, 2), (64, 64, 64), (16, 32, 128), (12, 10, 14), (128, 256, 64)])
def test_fft3d_complex(jt, prec, dims):
    pc.fftnd_complex(jt, prec, dims)


def test_fft3d_real(jt):
    pc.fftnd_real(jt, "Double", (32, 64, 16))
    pc.fftnd_real(jt, "Double", (2, 4, 8))
    pc.fftnd_real_full(jt, "Double", (8, 16, 32))
    pc.fftnd_real_full(jt, "Float", (6, 10, 12))


@pytest.mark.parametrize("prec,logn", [("Double", 22), ("Float", 23), ("Double", 24)])
def test_fft1d_three_pass_against_oracle(jt, prec, logn):
    """2^22 .. 2^26: three-pass composition of the lean kernels"""
    pc.fft1d_complex(jt, prec, 1 << logn)


def test_fft1d_2p27_three_pass_properties(jt):
    """n = 2^27 (2 GiB) is beyond the two-pass limit: three-pass path; Parseval, spot bins, round trip"""
    import ctypes
    import torch
    n = 1 << 27
    lib = __import__("jtransforms_b200")._lib.get()
    a = torch.empty(2 * n, dtype=torch.float64, device="cuda:0")
    assert lib.jtb_fill_uniform_device(0, 0, ctypes.c_void_p(a.data_ptr()), 2 * n, 7, -1.0, 1.0, None) == 0
    torch.cuda.synchronize()
    x = a.clone()
    f = jt.DoubleFFT_1D(n)
    f.complexForward(a)
    torch.cuda.synchronize()
    e_in, e_out = float((x * x).sum()), float((a * a).sum())
    assert abs(e_out / (n * e_in) - 1.0) < 1e-12
    xc = torch.view_as_complex(x.view(-1, 2))
    ac = torch.view_as_complex(a.view(-1, 2))
    j = torch.arange(n, device="cuda:0", dtype=torch.int64)
    for k in (0, 1, 12345, n // 2 + 3, n - 1):
        ph = ((j * k) % n).to(torch.float64) * (-2.0 * np.pi / n)      # exact phase reduction in integers
        want = torch.sum(xc * torch.polar(torch.ones_like(ph), ph))
        assert abs(complex(ac[k] - want)) <= 1e-12 * 27 * float(np.sqrt(e_in)) * 10, k
        del ph
    f.complexInverse(a, True)
    torch.cuda.synchronize()
    assert float(torch.linalg.norm(a - x) / torch.linalg.norm(x)) <= 1e-12 * 27


def test_fft3d_512_properties(jt):
    """config 5 at full size, device resident: spot bins against direct sums, Parseval, round trip"""
    import torch
    S = R = C = 512
    N = S * R * C
    lib = __import__("jtransforms_b200")._lib.get()
    a = torch.empty(2 * N, dtype=torch.float64, device="cuda:0")
    import ctypes
    assert lib.jtb_fill_uniform_device(0, 0, ctypes.c_void_p(a.data_ptr()), 2 * N, 2, 0.0, 1.0, None) == 0
    torch.cuda.synchronize()
    x = a.clone()
    # the device fill equals the oracle's counter-based generator bit for bit
    assert np.array_equal(x[:4096].cpu().numpy(), o.fill_uniform(4096, seed=2))
    f = jt.DoubleFFT_3D(S, R, C)
    f.complexForward(a)
    torch.cuda.synchronize()
    e_in = float((x * x).sum())
    e_out = float((a * a).sum())
    assert abs(e_out / (N * e_in) - 1.0) < 1e-12
    xc = torch.view_as_complex(x.view(-1, 2)).view(S, R, C)
    ac = torch.view_as_complex(a.view(-1, 2)).view(S, R, C)
    rng = np.random.default_rng(5)
    for _ in range(4):
        k1, k2, k3 = (int(v) for v in rng.integers(0, 512, 3))
        w1 = torch.exp(-2j * np.pi * k1 * torch.arange(S, device="cuda:0", dtype=torch.float64) / S)
        w2 = torch.exp(-2j * np.pi * k2 * torch.arange(R, device="cuda:0", dtype=torch.float64) / R)
        w3 = torch.exp(-2j * np.pi * k3 * torch.arange(C, device="cuda:0", dtype=torch.float64) / C)
        want = torch.einsum("srt,s,r,t->", xc, w1, w2, w3)
        got = ac[k1, k2, k3]
        assert abs(complex(got - want)) <= 1e-12 * 27 * float(torch.sqrt(torch.tensor(e_in * N) / N)) * 10
    f.complexInverse(a, True)
    torch.cuda.synchronize()
    err = float(torch.linalg.norm(a - x) / torch.linalg.norm(x))
    assert err <= 1e-12 * 27


@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
@pytest.mark.parametrize("dims", [(2,), (16,), (30,), (1024,), (8192,), (65536,), (1000,), (256, 256), (100, 60),
                                  (2048, 1024), (16, 32, 64), (5, 6, 7)])
def test_r2r(jt, kind, dims):
    pc.r2r(jt, "Double", kind, dims)


@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
def test_r2r_float(jt, kind):
    pc.r2r(jt, "Float", kind, (4096,))
    pc.r2r(jt, "Float", kind, (128, 512))


@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
@pytest.mark.parametrize("dims", [(4096, 4096), (8192, 512), (4, 4096, 128), (4096, 2, 64)])
def test_r2r_fused_paths(jt, kind, dims):
    pc.r2r(jt, "Double", kind, dims)


@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
@pytest.mark.parametrize("prec,dims", [("Double", (64, 64)), ("Double", (512, 512)), ("Double", (1024, 1024)), ("Double", (128, 48)),
                                       ("Double", (2048, 2048)), ("Float", (1024, 1024)), ("Float", (256, 64)),
                                       ("Double", (64, 64, 64)), ("Double", (32, 256, 128)), ("Double", (16384, 64))])
def test_r2r_single_pass_columns_and_inverse(jt, kind, prec, dims):
    """fused inverse kernels + single-pass column kernels (jtb_r2r_inv.cuh)"""
    pc.r2r(jt, prec, kind, dims)


def test_r2r_fast_inverse_is_taken(jt):
    from jtransforms_b200 import _lib
    L = _lib.get()
    for dims, want in (((1024, 1024), 2), ((8192, 8192), 3)):
        n = dims[0] * dims[1]
        x = o.fill_uniform(n, seed=9, lo=-1.0, hi=1.0)
        t = jt.DoubleDCT_2D(*dims)
        a = x.copy()
        t.inverse(a, True)          # first call builds tables
        a = x.copy()
        c0 = L.jtb_launch_count(0)
        t.inverse(a, True)
        assert L.jtb_launch_count(0) - c0 == want
        t.forward(a, True)
        assert o.rel_l2(a, x) < 1e-12 * 26


@pytest.mark.parametrize("offa", [1, 2, 3, 4])
def test_device_tensor_offsets_fall_back_cleanly(jt, offa):
    """device-resident arrays at offsets that break the 16/32-byte alignment the vectorised row kernels need"""
    import torch
    n = 1024
    x = o.fill_uniform(n + 8, seed=4, lo=-1.0, hi=1.0)
    for name, fn, want in (
            ("dct fwd", lambda t: jt.DoubleDCT_1D(n).forward(t, offa, True), lambda v: o.dct_forward_1d(v, True)),
            ("dct inv", lambda t: jt.DoubleDCT_1D(n).inverse(t, offa, True), lambda v: o.dct_inverse_1d(v, True)),
            ("dst inv", lambda t: jt.DoubleDST_1D(n).inverse(t, offa, False), lambda v: o.dst_inverse_1d(v, False)),
            ("real fwd", lambda t: jt.DoubleFFT_1D(n).realForward(t, offa), lambda v: o.real_forward_1d(v, n)),
            ("real inv", lambda t: jt.DoubleFFT_1D(n).realInverse(t, offa, True), lambda v: o.real_inverse_1d(v, n, True))):
        t = torch.from_numpy(x.copy()).cuda()
        fn(t)
        got = t.cpu().numpy()
        assert np.array_equal(got[:offa], x[:offa]) and np.array_equal(got[offa + n:], x[offa + n:]), name
        assert o.rel_l2(got[offa:offa + n], want(x[offa:offa + n])) < 1e-12 * 10, name


@pytest.mark.parametrize("prec,dims", [("Double", (512, 1024)), ("Float", (256, 128)), ("Double", (64, 128, 256)),
                                       ("Double", (2, 4096)), ("Double", (4096, 2))])
def test_real_full_packed_plus_expansion(jt, prec, dims):
    pc.fftnd_real_full(jt, prec, dims)


def test_staged_pageable_copies(jt):
    """plain (pageable) numpy arrays above 32 MiB go through the multi-threaded staging path; pinned ones do not"""
    n = 1 << 22                                   # 64 MiB of complex doubles
    x = o.fill_uniform(2 * n, seed=12, lo=-1.0, hi=1.0)
    a = x.copy()
    f = jt.DoubleFFT_1D(n)
    f.complexForward(a)
    want = o.complex_forward_1d(x, n)
    assert o.rel_l2(a, want) < 1e-12 * 22
    f.complexInverse(a, True)
    assert o.rel_l2(a, x) < 1e-12 * 22
    # odd byte count / ragged last chunk, float, real-full (input span = half of the output span)
    pc.fft1d_batch(jt, "Float", 1000003, 5, pad=2)
    pc.fftnd_real_full(jt, "Double", (2048, 4096))


def test_host_register_roundtrip(jt):
    """jtb_host_register / jtb_host_unregister around a caller-owned array (pinned context manager)"""
    from jtransforms_b200.utils import pinned
    n = 1 << 18
    x = o.fill_uniform(2 * n, seed=8, lo=-1.0, hi=1.0)
    a = x.copy()
    with pinned(a):
        jt.DoubleFFT_1D(n).complexForward(a)
        jt.DoubleFFT_1D(n).complexInverse(a, True)
    assert o.rel_l2(a, x) < 1e-12 * 18


def test_fft1d_batch_pipelined(jt, monkeypatch):
    """jtb_exec_batch as a three-stage pipeline over chunks of transforms (default 64 MiB chunks; here 1 MiB)"""
    monkeypatch.setenv("JTB_BATCH_MB", "1")
    pc.fft1d_batch(jt, "Double", 4096, 100)           # 64 KiB per transform -> 16 per chunk, ragged tail
    pc.fft1d_batch(jt, "Float", 1000, 700, pad=6)     # Bluestein-free mixed radix, padded distance
    pc.fft1d_batch(jt, "Float", 10007, 64)            # prime length (Bluestein) through the ring


def test_dct2d_8192(jt):
    """config 4 at full size (DCT; DST/DHT share every kernel and are covered at 2048x1024 above)"""
    import scipy.fft as sfft
    n = 8192
    x = o.fill_uniform(n * n, seed=2)
    a = x.copy()
    jt.DoubleDCT_2D(n, n).forward(a, True)
    want = sfft.dctn(x.reshape(n, n), type=2, norm="ortho", workers=-1).ravel()
    pc.check(a, want, "Double", n * n, "DCT 8192^2")


@pytest.mark.parametrize("kind", ["DST", "DHT"])
def test_dst_dht_2d_8192(jt, kind):
    """config 4 names all three transforms: DoubleDST_2D / DoubleDHT_2D forward at 8192 x 8192 against the oracle
    (dst/DoubleDST_2D.java:103, dht/DoubleDHT_2D.java:102-190 + yTransform :1288-1309), element for element"""
    n = 8192
    x = o.fill_uniform(n * n, seed=3, lo=-1.0, hi=1.0)
    a = x.copy()
    if kind == "DST":
        jt.DoubleDST_2D(n, n).forward(a, True)
        want = o.dst_forward_nd(x, (n, n), True)
    else:
        jt.DoubleDHT_2D(n, n).forward(a)
        want = o.dht_forward_nd(x, (n, n))
    pc.check(a, want, "Double", n * n, "%s 8192^2" % kind)


def test_fft3d_512_vs_oracle(jt):
    """config 5 itself, element for element: the host-array path (jtb_exec) at 512^3 against
    oracle.complex_forward_3d, tolerance 1e-12 * log2(N) = 27e-12 relative L2 -- and the same array through the
    8-way slab decomposition (single-process multi-GPU plan, eight virtual ranks on this GPU: the shapes, kernels and
    exchange addressing of the 8-GPU run)"""
    S = R = C = 512
    x = o.fill_uniform(2 * S * R * C, seed=2)
    want = o.complex_forward_3d(x, S, R, C)
    a = x.copy()
    jt.DoubleFFT_3D(S, R, C).complexForward(a)
    err = o.rel_l2(a, want)
    assert err <= 27e-12, err
    a[:] = x
    jt.DoubleFFT_3D(S, R, C, devices=[0] * 8).complexForward(a)
    err8 = o.rel_l2(a, want)
    assert err8 <= 27e-12, err8


def test_large_64bit_indexing(jt):
    """arrays beyond 2^31 bytes / 2^30 elements, device resident: 1024^3 complex double (16 GiB) round trip +
    Parseval, and a 2^26-point 1-D transform against the oracle"""
    import ctypes
    import torch
    lib = __import__("jtransforms_b200")._lib.get()
    S = R = C = 1024
    N = S * R * C
    a = torch.empty(2 * N, dtype=torch.float64, device="cuda:0")
    assert lib.jtb_fill_uniform_device(0, 0, ctypes.c_void_p(a.data_ptr()), 2 * N, 7, -1.0, 1.0, None) == 0
    torch.cuda.synchronize()
    e_in = float((a * a).sum())
    probe = a[-4096:].clone()
    f = jt.DoubleFFT_3D(S, R, C)
    f.complexForward(a)
    torch.cuda.synchronize()
    e_out = float((a * a).sum())
    assert abs(e_out / (N * e_in) - 1.0) < 1e-12
    f.complexInverse(a, True)
    torch.cuda.synchronize()
    assert float(torch.linalg.norm(a[-4096:] - probe) / torch.linalg.norm(probe)) < 1e-12 * 30
    del a
    torch.cuda.empty_cache()
    n = 1 << 26
    x = o.fill_uniform(2 * n, seed=13, lo=-1.0, hi=1.0)
    b = x.copy()
    jt.DoubleFFT_1D(n).complexForward(b)
    pc.check(b, o.complex_forward_1d(x, n), "Double", n, "2^26")


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("n", [10368, 27000, 75600, 165375, 362880, 1562500, 3211264, 6250000])
def test_fft1d_smooth_reference_sizes(jt, prec, n):
    """the reference benchmark's mixed-radix lengths (fft/BenchmarkDoubleFFT.java:56; FFTPACK path
    fft/DoubleFFT_1D.java:6630-8009): two mixed-radix passes n = N1*N2 here, against the oracle"""
    pc.fft1d_complex(jt, prec, n)


def test_host_array_beyond_2p31_elements(jt):
    """SURVEY 8(f) rank 4, the LargeArray range at the boundary: ONE host array of 2^32 floats (16 GiB, more elements
    than a Java array or a 32-bit index can hold -- what DoubleLargeArray / FloatLargeArray exist for,
    fft/DoubleFFT_2D.java:230) through jtb_exec: FloatFFT_2D 32768 x 65536 complexForward of two impulses, checked
    against the closed form at 200 000 sampled bins, at both ends of the array, and through Parseval."""
    import psutil
    if psutil.virtual_memory().available < (40 << 30):
        pytest.skip("needs 40 GiB of free host memory")
    R, C = 32768, 65536
    a = np.zeros(2 * R * C, dtype=np.float32)
    assert a.size == 1 << 32
    imp = [(5, 7, 1.5, -0.25), (R - 3, C - 11, -0.75, 2.0)]          # (row, column, re, im)
    for r0, c0, re, im in imp:
        a[2 * (r0 * C + c0)] = re
        a[2 * (r0 * C + c0) + 1] = im
    jt.FloatFFT_2D(R, C).complexForward(a)
    z = a.view(np.complex64)
    rng = np.random.default_rng(5)
    k1 = np.concatenate([rng.integers(0, R, 200000), [0, 0, R - 1, R - 1]])
    k2 = np.concatenate([rng.integers(0, C, 200000), [0, C - 1, 0, C - 1]])
    want = np.zeros(k1.size, dtype=np.complex128)
    for r0, c0, re, im in imp:
        want += (re + 1j * im) * np.exp(-2j * np.pi * ((k1 * r0) % R / R + (k2 * c0) % C / C))
    got = z[k1.astype(np.int64) * C + k2].astype(np.complex128)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 1e-5 * 31, err
    energy = 0.0
    for blk in np.array_split(np.arange(R), 64):                        # Parseval without a 16 GiB temporary
        v = a[2 * blk[0] * C:2 * (blk[-1] + 1) * C]
        energy += float(np.sum(np.square(v.astype(np.float64))))       # FP64 accumulation (a float32 dot loses 1e-3 here)
    expect = float(R) * C * sum(re * re + im * im for _, _, re, im in imp)
    assert abs(energy - expect) / expect < 1e-4, (energy, expect)
