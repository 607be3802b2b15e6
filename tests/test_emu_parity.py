"""CPU-only parity run of the library's REAL kernel sources through the g++ CUDA-semantics shim
(tests/emu).  This checks index math, packed layouts, fusions and host logic without a GPU; the GPU
suite (tests/test_gpu_parity.py) repeats the same cases through the nvcc-built libjtb200.so.
"""
import os
import subprocess

import numpy as np
import pytest

import parity_cases as pc
from oracle import jt_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_LIB = os.path.join(HERE, "emu", "_build", "libjtb200_emu.so")


@pytest.fixture(scope="module")
def jt():
    subprocess.run(["sh", os.path.join(HERE, "emu", "build_emu.sh")], check=True, capture_output=True)
    import jtransforms_b200 as m
    from jtransforms_b200 import _lib
    _lib.use(EMU_LIB)
    yield m
    _lib.get().jtb_debug_set_limits(0, 0)
    _lib._lib = None


@pytest.fixture
def small_limits(jt):
    """force the two-pass (four-step) path at sizes the emulator handles quickly"""
    from jtransforms_b200 import _lib
    _lib.get().jtb_debug_set_limits(5, 3)
    yield
    _lib.get().jtb_debug_set_limits(0, 0)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 12, 13, 16, 32, 64, 100, 120, 128, 211, 256, 310, 512])
def test_fft1d_complex(jt, prec, n):
    pc.fft1d_complex(jt, prec, n)


def test_fft1d_offset(jt):
    pc.fft1d_complex(jt, "Double", 64, offa=6)
    pc.fft1d_complex(jt, "Double", 30, offa=3)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 9, 16, 30, 64, 101, 128, 256])
def test_fft1d_real(jt, prec, n):
    pc.fft1d_real(jt, prec, n)


@pytest.mark.parametrize("n", [64, 128, 256, 1024])
def test_fft1d_two_pass(jt, small_limits, n):
    pc.fft1d_complex(jt, "Double", n)
    pc.fft1d_batch(jt, "Float", n, 3, pad=4)


@pytest.mark.parametrize("n", [100, 311])
def test_fft1d_bluestein_two_pass(jt, small_limits, n):
    pc.fft1d_complex(jt, "Double", n)
    pc.fft1d_real(jt, "Double", n)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("dims", [(2, 2), (2, 8), (8, 2), (16, 4), (64, 32), (2, 2, 2), (2, 4, 8), (8, 2, 4), (4, 8, 2),
                                  (8, 16, 32)])
def test_real_full_packed_plus_expansion(jt, prec, dims):
    """realForwardFull / realInverseFull of power-of-two sizes: packed transform + Hermitian expansion kernel"""
    pc.fftnd_real_full(jt, prec, dims)


@pytest.mark.parametrize("n", [4, 8, 64, 1024])
def test_real_full_1d_expansion(jt, n):
    x = o.fill_uniform(n, seed=3, lo=-1.0, hi=1.0)
    for inverse in (False, True):
        a = np.zeros(2 * n)
        a[:n] = x
        z = a.copy()
        if inverse:
            jt.DoubleFFT_1D(n).realInverseFull(a, True)
            want = o.real_inverse_full_1d(z, n, True)
        else:
            jt.DoubleFFT_1D(n).realForwardFull(a)
            want = o.real_forward_full_1d(z, n)
        assert o.rel_l2(a, want) < 1e-12 * 12


def test_fft1d_three_pass(jt):
    """lines beyond the two-pass limit (Engine::c2c_big_contig), exercised at small sizes with the limits forced down"""
    import sys
    env = dict(os.environ, JTB_NO_FAST="1", JTB_NO_FAST2="1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "three_pass_emu.py"), EMU_LIB], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "three-pass ok" in r.stdout, r.stdout + r.stderr


def test_fft1d_three_pass_lean(jt):
    """fast_threepass_contig: strided two-pass sub-transform with the outer twiddle fused + transposing row pass"""
    import sys
    env = dict(os.environ, JTB_THREEPASS_MIN="19")
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "three_pass_lean_emu.py"), EMU_LIB], env=env,
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and "three-pass lean ok" in r.stdout, r.stdout + r.stderr


def test_host_register_symbols(jt):
    from jtransforms_b200.utils import pinned
    a = o.fill_uniform(128, seed=1)
    x = a.copy()
    with pinned(a):
        jt.DoubleFFT_1D(64).complexForward(a)
    assert o.rel_l2(a, o.complex_forward_1d(x, 64)) < 1e-12 * 6


def test_staged_pageable_copies(jt, monkeypatch):
    """pageable caller memory: multi-threaded staging through bounce buffers (jtb_stage.cu), several 8 MiB chunks,
    ragged tail, real-full op whose input is half of the output span"""
    monkeypatch.setenv("JTB_EMU_PAGEABLE", "1")
    monkeypatch.setenv("JTB_STAGE_MIN_MB", "0")
    monkeypatch.setenv("JTB_STAGE_THREADS", "3")
    pc.fft1d_batch(jt, "Double", 1024, 1100, pad=2)       # 17.2 MiB -> 3 chunks
    pc.fft1d_complex(jt, "Double", 64)
    pc.fftnd_real_full(jt, "Double", (16, 8))


def test_fft1d_batch_pipelined(jt, monkeypatch):
    """jtb_exec_batch in chunks (three-slot H2D / kernels / D2H ring): ragged last chunk, padded distance"""
    monkeypatch.setenv("JTB_BATCH_MB", "0.004")       # 4 KiB chunks: 64-point double transforms -> 4 per chunk
    pc.fft1d_batch(jt, "Double", 64, 23)
    pc.fft1d_batch(jt, "Float", 100, 17, pad=6)
    pc.fft1d_batch(jt, "Double", 31, 40, pad=2)


def test_fft1d_batch(jt):
    pc.fft1d_batch(jt, "Double", 64, 5)
    pc.fft1d_batch(jt, "Float", 17, 4, pad=2)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("dims", [(2, 2), (4, 8), (16, 4), (5, 6), (12, 7), (32, 32), (3, 16)])
def test_fft2d_complex(jt, prec, dims):
    pc.fftnd_complex(jt, prec, dims)


@pytest.mark.parametrize("dims", [(2, 2), (2, 8), (8, 2), (4, 16), (16, 16), (32, 8)])
def test_fft2d_real(jt, dims):
    pc.fftnd_real(jt, "Double", dims)
    pc.fftnd_real_full(jt, "Double", dims)


def test_fft2d_real_full_nonpow2(jt):
    pc.fftnd_real_full(jt, "Double", (6, 10))


def test_fft2d_two_pass(jt, small_limits):
    pc.fftnd_complex(jt, "Double", (64, 16))
    pc.fftnd_complex(jt, "Double", (16, 128))
    pc.fftnd_real(jt, "Double", (32, 64))


@pytest.mark.parametrize("dims", [(2, 2, 2), (4, 4, 8), (8, 2, 4), (3, 5, 4), (16, 8, 4)])
def test_fft3d_complex(jt, dims):
    pc.fftnd_complex(jt, "Double", dims)


@pytest.mark.parametrize("dims", [(2, 2, 2), (4, 4, 8), (8, 4, 2), (2, 8, 4), (8, 8, 8)])
def test_fft3d_real(jt, dims):
    pc.fftnd_real(jt, "Double", dims)
    pc.fftnd_real_full(jt, "Double", dims)


def test_fft3d_two_pass(jt, small_limits):
    pc.fftnd_complex(jt, "Double", (32, 16, 8))


@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
@pytest.mark.parametrize("dims", [(2,), (8,), (16,), (9,), (30,), (64,), (4, 8), (6, 5), (16, 16), (4, 2, 8), (3, 4, 5)])
def test_r2r(jt, kind, dims):
    pc.r2r(jt, "Double", kind, dims)


@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
def test_r2r_float(jt, kind):
    pc.r2r(jt, "Float", kind, (32,))
    pc.r2r(jt, "Float", kind, (8, 16))


def test_errors(jt):
    with pytest.raises(ValueError, match="greater than 0"):
        jt.DoubleFFT_1D(0)
    with pytest.raises(ValueError, match="greater than 1"):
        jt.DoubleFFT_2D(1, 8)
    with pytest.raises(ValueError, match="greater than 1"):
        jt.DoubleFFT_3D(4, 1, 8)
    with pytest.raises(ValueError, match="power of two"):
        jt.DoubleFFT_2D(6, 8).realForward(np.zeros(48))
    with pytest.raises(ValueError, match="power of two"):
        jt.DoubleFFT_3D(4, 6, 8).realInverse(np.zeros(4 * 6 * 8), True)
    with pytest.raises(IndexError):
        jt.DoubleFFT_1D(16).complexForward(np.zeros(16))
    with pytest.raises(ValueError):
        jt.DoubleFFT_1D(16).complexForward(np.zeros(32, dtype=np.float32))
    a = np.array([3.0, 4.0])
    jt.DoubleFFT_1D(1).complexForward(a)       # n == 1 is a no-op
    assert a.tolist() == [3.0, 4.0]


# shapes that dispatch to the lean fft_fast_kernel (jtb_fast.cuh): contiguous and strided layouts
@pytest.mark.parametrize("prec,dims", [("Double", (512, 16)), ("Double", (64, 64)), ("Double", (1024, 8)),
                                       ("Double", (8, 2048)), ("Double", (2, 4096)), ("Float", (512, 32)),
                                       ("Float", (1024, 16)), ("Float", (2048, 8)), ("Double", (8, 16, 512))])
def test_fast_kernel_shapes(jt, prec, dims):
    pc.fftnd_complex(jt, prec, dims)


def test_fast_kernel_variants(jt, monkeypatch):
    for ws, wc in ((4, 1), (8, 2), (4, 8)):
        monkeypatch.setenv("JTB_FAST_WS", str(ws))
        monkeypatch.setenv("JTB_FAST_WC", str(wc))
        pc.fftnd_complex(jt, "Double", (512, 8) if ws <= 8 else (512, 16))
        pc.fft1d_batch(jt, "Double", 512, 3, pad=2)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_slab_round_trip_virtual_ranks(jt, P):
    """64 slices: the way back (jtb_fft3d_k1_scatter, inverse re-slabbing fused into the k1 pass) is exercised too"""
    from jtransforms_b200 import _lib
    pc.slab_scatter_virtual(_lib.get(), "Double", (64, 64, 8), P)
    if P == 2:
        pc.slab_scatter_virtual(_lib.get(), "Float", (64, 64, 16), P)


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_slab_scatter_virtual_ranks(jt, P):
    from jtransforms_b200 import _lib
    pc.slab_scatter_virtual(_lib.get(), "Double", (8, 64, 16), P)


def test_slab_scatter_float(jt):
    from jtransforms_b200 import _lib
    pc.slab_scatter_virtual(_lib.get(), "Float", (4, 64, 32), 2)


# lean two-pass paths (jtb_fast2.cuh) at their natural sizes
@pytest.mark.parametrize("n", [1 << 14, 1 << 16, 1 << 17])
def test_fast2_fourstep_contig(jt, n):
    pc.fft1d_complex(jt, "Double", n)


def test_fast2_fourstep_contig_batch_float(jt):
    pc.fft1d_batch(jt, "Float", 1 << 18, 2, pad=2)


@pytest.mark.parametrize("dims", [(4096, 32), (8192, 32)])
def test_fast2_fourstep_strided(jt, dims):
    pc.fftnd_complex(jt, "Double", dims)


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
def test_fast2_rfft(jt, n):
    pc.fft1d_real(jt, "Double", n)


def test_fast2_real2d(jt):
    pc.fftnd_real(jt, "Double", (8, 4096))
    pc.fftnd_real(jt, "Float", (4, 2048))


def test_fast2_fourstep_strided_strips(jt, monkeypatch):
    monkeypatch.setenv("JTB_STRIP_MB", "2")       # 4 column strips of 32 for 4096 x 128
    pc.fftnd_complex(jt, "Double", (4096, 128))


# fused forward DCT/DST/DHT kernels (rows: fft_r2r_row_kernel; columns: permuted two-pass + pair post-pass)
@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
@pytest.mark.parametrize("dims", [(512,), (4, 1024), (4096, 64)])
def test_fast2_r2r(jt, kind, dims):
    pc.r2r(jt, "Double", kind, dims)


@pytest.mark.parametrize("dims", [(8192,), (8192, 32)])
def test_fast2_r2r_long(jt, dims):
    pc.r2r(jt, "Double", "DCT", dims)


# fused inverse kernels and the single-pass column kernels (jtb_r2r_inv.cuh): every length class
@pytest.mark.parametrize("kind", ["DCT", "DST", "DHT"])
@pytest.mark.parametrize("prec,dims", [("Double", (64, 64)), ("Double", (256, 32)), ("Double", (128, 48)), ("Double", (1024, 16)),
                                       ("Double", (32, 16)), ("Double", (64, 48)),
                                       ("Float", (64, 32)), ("Float", (512, 32)), ("Double", (32, 32, 32)),
                                       ("Double", (2048, 64))])
def test_r2r_single_pass_columns_and_inverse(jt, kind, prec, dims):
    pc.r2r(jt, prec, kind, dims)


def test_r2r_fast_inverse_is_taken(jt):
    """the inverse of a fused size is one launch per axis (rows) / one or two (columns), not the staged pipeline"""
    from jtransforms_b200 import _lib
    L = _lib.get()
    x = o.fill_uniform(256 * 64, seed=9, lo=-1.0, hi=1.0)
    t = jt.DoubleDCT_2D(256, 64)
    a = x.copy()
    c0 = L.jtb_launch_count(0)
    t.inverse(a, True)
    assert L.jtb_launch_count(0) - c0 == 2
    assert o.rel_l2(a, o.dct_inverse_nd(x, (256, 64), True)) < 1e-12 * 14
    c0 = L.jtb_launch_count(0)
    t.forward(a, True)
    assert L.jtb_launch_count(0) - c0 == 2
    assert o.rel_l2(a, x) < 1e-12 * 14


def test_fft2d_2048_rows_two_pass_columns(jt):
    """2048-point strided lines: 64 x 32 two-pass split"""
    pc.fftnd_complex(jt, "Double", (2048, 32))


def test_fast2_r2r_strips_and_float(jt, monkeypatch):
    monkeypatch.setenv("JTB_STRIP_MB", "1")
    pc.r2r(jt, "Double", "DCT", (4096, 128))
    pc.r2r(jt, "Float", "DST", (2, 2048))


def test_fast_bluestein(jt):
    pc.fft1d_complex(jt, "Double", 140001)
    pc.fft1d_batch(jt, "Float", 131073, 2, pad=2)


@pytest.mark.parametrize("prec,dims", [("Double", (256, 16)), ("Double", (128, 32)), ("Float", (256, 32)),
                                       ("Float", (128, 64)), ("Float", (64, 64)), ("Double", (16, 256)),
                                       ("Float", (32, 128))])
def test_fast_kernel_small_sizes(jt, prec, dims):
    pc.fftnd_complex(jt, prec, dims)


def test_fast_kernel_8192_rows(jt):
    pc.fft1d_batch(jt, "Double", 8192, 2, pad=2)
    pc.fft1d_batch(jt, "Float", 8192, 2)
    pc.fft1d_complex(jt, "Float", 4096)


# native mixed-radix lengths (jtb_mixed.cuh): 2^a 3^b 5^c 7^d 11^e 13^f that fit one CTA
@pytest.mark.parametrize("n", [3, 5, 6, 7, 9, 10, 11, 12, 13, 15, 100, 120, 1056, 420, 1000, 1050, 2916])
def test_mixed_radix_1d(jt, n):
    pc.fft1d_complex(jt, "Double", n)


def test_mixed_radix_float_real_nd(jt):
    pc.fft1d_complex(jt, "Float", 360)
    pc.fft1d_real(jt, "Double", 300)
    pc.fft1d_real(jt, "Double", 45)
    pc.fftnd_complex(jt, "Double", (30, 42))
    pc.fftnd_complex(jt, "Double", (12, 10, 14))
    pc.fftnd_complex(jt, "Float", (24, 36))
    pc.r2r(jt, "Double", "DCT", (30, 20))
    pc.fft1d_batch(jt, "Double", 100, 7, pad=2)


def test_slices_entry_point_virtual_ranks(jt):
    """jtb_fft2d_slices_device with receive buffers (general path in the emulated build)"""
    from jtransforms_b200 import _lib
    pc.slab_scatter_virtual(_lib.get(), "Double", (4, 64, 64), 2, fused_slices=True)


# ------------------------------------------------------------------ multi-GPU plans (virtual ranks on the emulated device)
@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("P,dims", [(2, (8, 8, 16)), (4, (8, 8, 16)), (2, (6, 10, 5)), (8, (64, 64, 16)), (4, (4, 64, 64)),
                                    (3, (6, 9, 4))])
def test_multi_device_plan_virtual(jt, prec, P, dims):
    """jtb_plan_set_devices: ONE host array in, natural order out, slab-decomposed over P members (same device listed
    P times); forward against the oracle, then complexInverse(scale) back to the input"""
    pc.fft3d_multi(jt, prec, dims, [0] * P)


def test_multi_device_plan_pageable(jt, monkeypatch):
    """pageable caller memory: the per-member staged (pitched) copies deliver the natural order too"""
    monkeypatch.setenv("JTB_EMU_PAGEABLE", "1")
    monkeypatch.setenv("JTB_STAGE_MIN_MB", "0")
    pc.fft3d_multi(jt, "Double", (8, 8, 16), [0, 0])
    pc.fft3d_multi(jt, "Double", (6, 10, 5), [0, 0])


def test_multi_device_batch(jt):
    """batches on a multi-GPU plan are split into one contiguous block per member"""
    n, howmany = 100, 7
    x = pc.rnd(2 * n * howmany)
    a = x.copy()
    jt.DoubleFFT_1D(n, devices=[0, 0, 0]).complexForwardBatch(a, howmany, 2 * n)
    for b in range(howmany):
        pc.check(a[2 * n * b:2 * n * (b + 1)], o.complex_forward_1d(x[2 * n * b:2 * n * (b + 1)], n), "Double", n, "batch %d" % b)
    # other ops of a multi-GPU plan run on devices[0]
    y = pc.rnd(8 * 8 * 16)
    b2 = y.copy()
    jt.DoubleFFT_3D(8, 8, 16, devices=[0, 0]).realForward(b2)
    pc.check(b2, o.real_forward_3d(y, 8, 8, 16), "Double", 1024, "realForward on a multi plan")


def test_multi_device_errors(jt):
    from jtransforms_b200 import _lib
    with pytest.raises(Exception):
        jt.DoubleFFT_3D(8, 8, 16, devices=[0, 7])          # no such device
    f = jt.DoubleFFT_3D(6, 8, 16, devices=[0, 0, 0, 0])      # 6 slices do not divide by 4: no slab path, devices[0] runs it
    x = pc.rnd(2 * 6 * 8 * 16)
    a = x.copy()
    f.complexForward(a)
    pc.check(a, o.complex_forward_3d(x, 6, 8, 16), "Double", 768, "non-divisible falls back to one device")
    assert _lib.get().jtb_plan_device_count(f._plan._h) == 4


def test_exec_n_bounds(jt):
    """the C entry point refuses arrays shorter than offa + elements (the Java shim's ArrayIndexOutOfBounds)"""
    import ctypes as C
    from jtransforms_b200 import _lib
    lib = _lib.get()
    f = jt.DoubleFFT_1D(64)
    a = np.zeros(128)
    assert lib.jtb_exec_n(f._plan._h, _lib.C2C_FORWARD, C.c_void_p(a.ctypes.data), 127, 0, 0) == _lib.ERR_ARG
    assert b"too short" in lib.jtb_last_error()
    assert lib.jtb_exec_n(f._plan._h, _lib.C2C_FORWARD, C.c_void_p(a.ctypes.data), 128, 1, 0) == _lib.ERR_ARG
    assert lib.jtb_exec_n(f._plan._h, _lib.C2C_FORWARD, C.c_void_p(a.ctypes.data), 128, 0, 0) == _lib.OK


def test_plan_destroy_releases_tables(jt):
    """a caller sweeping over sizes must not grow the table cache: the chirp tables of a Bluestein plan go with it"""
    from jtransforms_b200 import _lib
    lib = _lib.get()
    base = lib.jtb_debug_table_bytes(0)
    for n in (10007, 10009, 10037):
        f = jt.DoubleFFT_1D(n)
        a = pc.rnd(2 * n)
        f.complexForward(a)
        during = lib.jtb_debug_table_bytes(0)
        assert during > base + 16 * n
        del f
        import gc
        gc.collect()
        assert lib.jtb_debug_table_bytes(0) < base + 64 * 1024 * 8, "tables of a destroyed plan were not released"
    # a table shared by two plans survives the first destroy and still gives right answers
    f1, f2 = jt.DoubleFFT_1D(10007), jt.DoubleFFT_1D(10007)
    x = pc.rnd(2 * 10007)
    a1 = x.copy(); f1.complexForward(a1)
    a2 = x.copy(); f2.complexForward(a2)
    del f1
    gc.collect()
    a3 = x.copy(); f2.complexForward(a3)
    assert np.array_equal(a2, a3)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("n", [10368, 27000, 9 * 1024, 3 * 5 * 7 * 11 * 13, 20000])
def test_fft1d_mixed_two_pass(jt, prec, n):
    """smooth lengths beyond one CTA: two mixed-radix passes n = N1*N2 with the four-step twiddle fused into the
    first store and a transposed second store (the reference's FFTPACK sizes, fft/BenchmarkDoubleFFT.java:56)"""
    pc.fft1d_complex(jt, prec, n)


def test_fft1d_mixed_two_pass_batch_and_real(jt):
    """batched lines (jtb_exec_batch) and the real transforms that sit on top of the complex one"""
    n = 10368
    x = o.fill_uniform(2 * n * 3, seed=4, lo=-1.0, hi=1.0)
    a = x.copy()
    jt.DoubleFFT_1D(n).complexForwardBatch(a, 3, 2 * n)
    for b in range(3):
        assert o.rel_l2(a[2 * n * b:2 * n * (b + 1)], o.complex_forward_1d(x[2 * n * b:2 * n * (b + 1)], n)) < 1e-12 * 14
    pc.fft1d_real(jt, "Double", 12000)


@pytest.mark.parametrize("chunks", ["2", "4"])
def test_multi_device_plan_pipelined_exchange(jt, chunks):
    """the opt-in column-block pipelined exchange (JTB_SLAB_CHUNKS: windowed scatter launches, per-block events,
    slice-axis pass per block on the second stream) gives the same result as the default step"""
    import sys
    env = dict(os.environ, JTB_SLAB_CHUNKS=chunks)
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "pipe_group_emu.py"), EMU_LIB], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("dims", [(64, 128), (2, 64), (128, 1024), (16, 8192)])
def test_dht2d_rows_with_folded_ytransform(jt, prec, dims):
    """DoubleDHT_2D: the row pass owns the row pairs (r, R-r) and applies yTransform (dht/DoubleDHT_2D.java:1288-1309) in
    its store (fast_dht2d_rows) -- forward, scaled and unscaled inverse against the oracle"""
    pc.r2r(jt, prec, "DHT", dims)
    from jtransforms_b200 import _lib
    lib = _lib.get()
    x = pc.rnd(dims[0] * dims[1]).astype(pc.dtype_of(prec))
    l0 = lib.jtb_launch_count(0)
    getattr(jt, prec + "DHT_2D")(*dims).forward(x)
    assert lib.jtb_launch_count(0) - l0 <= (2 if dims[0] >= 32 else 4)      # no separate yTransform launch


@pytest.mark.parametrize("prec", ["Float", "Double"])
def test_batch_split_over_plan_devices(jt, prec):
    """jtb_exec_batch on a multi-GPU plan: contiguous blocks of the batch per listed device (ragged: 7 lines over 3
    members), one host thread and one pipeline each -- BASELINE config 3's sharding behind the drop-in call"""
    n, howmany = 1009, 7
    dt = pc.dtype_of(prec)
    x = pc.rnd(2 * n * howmany).astype(dt)
    a = x.copy()
    f = getattr(jt, prec + "FFT_1D")(n, devices=[0, 0, 0])
    f.complexForwardBatch(a, howmany, 2 * n)
    for b in range(howmany):
        want = o.complex_forward_1d(x[2 * n * b:2 * n * (b + 1)].astype(np.float64), n)
        pc.check(a[2 * n * b:2 * n * (b + 1)], want, prec, n, "batch line %d" % b)
