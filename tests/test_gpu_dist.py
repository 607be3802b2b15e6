"""Multi-GPU tests (need >= 2 GPUs; skipped otherwise) and single-GPU virtual-rank tests of the fused exchange."""
import os
import subprocess
import sys

import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("P,dims", [(8, (64, 512, 64)), (2, (16, 512, 128)), (4, (32, 64, 40)), (1, (8, 1024, 16))])
def test_slab_scatter_virtual_ranks(P, dims):
    from jtransforms_b200 import _lib
    _lib._lib = None
    pc.slab_scatter_virtual(_lib.get(), "Double", dims, P, torch_device="cuda:0")


@pytest.mark.parametrize("P,dims", [(8, (16, 512, 512)), (2, (8, 512, 512)), (4, (8, 64, 64))])
def test_fused_slices_scatter_virtual_ranks(P, dims):
    """jtb_fft2d_slices_device with receive buffers: for 512 x 512 double slices this is the persistent kernel that
    fuses rows, columns and the exchange (the default multi-GPU path)"""
    from jtransforms_b200 import _lib
    _lib._lib = None
    pc.slab_scatter_virtual(_lib.get(), "Double", dims, P, torch_device="cuda:0", fused_slices=True)


@pytest.mark.parametrize("prec", ["Double", "Float"])
@pytest.mark.parametrize("P,dims", [(2, (64, 64, 64)), (4, (16, 512, 512)), (8, (64, 512, 64)), (2, (6, 10, 5)), (4, (128, 128, 128)),
                                    (8, (256, 256, 32)), (3, (6, 9, 4))])
def test_multi_device_plan_virtual_ranks(prec, P, dims):
    """jtb_plan_set_devices + jtb_exec through the host API with the SAME GPU listed P times: the single-process
    slab decomposition (H2D per member, fused passes + exchange, natural-order pitched D2H) on a 1-GPU box"""
    import jtransforms_b200 as jt
    from jtransforms_b200 import _lib
    _lib._lib = None
    pc.fft3d_multi(jt, prec, dims, [0] * P)


def test_multi_device_plan_pinned_and_pageable():
    """pinned (jtb_host_alloc) and pageable caller arrays take different copy paths; both must deliver natural order"""
    import ctypes as C
    import numpy as np
    import jtransforms_b200 as jt
    from jtransforms_b200 import _lib
    from oracle import jt_oracle as o
    _lib._lib = None
    lib = _lib.get()
    S, R, Cn = 64, 128, 256            # 32 MiB: above the staging threshold
    x = o.fill_uniform(2 * S * R * Cn, seed=5, lo=-1.0, hi=1.0)
    want = o.complex_forward_3d(x, S, R, Cn)
    f = jt.DoubleFFT_3D(S, R, Cn, devices=[0, 0, 0, 0])
    a = x.copy()                       # pageable
    f.complexForward(a)
    assert o.rel_l2(a, want) < 1e-12 * 21
    hp = C.c_void_p()
    _lib.check(lib.jtb_host_alloc(C.byref(hp), x.nbytes))
    b = np.ctypeslib.as_array((C.c_double * x.size).from_address(hp.value))
    b[:] = x
    f.complexForward(b)
    assert o.rel_l2(b, want) < 1e-12 * 21
    f.complexInverse(b, True)
    assert o.rel_l2(b, x) < 1e-12 * 21
    del b
    lib.jtb_host_free(hp)


def test_multi_device_plan_all_gpus():
    """the real thing: one host array over every GPU of the box, peer stores and (second run) the library's NCCL exchange"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import jtransforms_b200 as jt, parity_cases as pc\n"
            "for dims in [(64, 64, 64), (2 * %d, 512, 512), (128, 128, 128), (6 * %d, 5 * %d, 10)]:\n"
            "    pc.fft3d_multi(jt, 'Double', dims, list(range(%d)))\n"
            "pc.fft3d_multi(jt, 'Float', (64, 512, 64), list(range(%d)))\n"
            "print('ok')\n") % (os.path.dirname(HERE), HERE, n, n, n, n, n)
    for extra in ({}, {"JTB_EXCHANGE_NCCL": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, str(extra) + r.stdout[-3000:] + r.stderr[-3000:]


def test_slab_two_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(HERE, "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_plans_on_two_devices_in_one_process():
    """per-device state (tables, shared-memory attributes, streams) must not leak between devices"""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import jtransforms_b200 as jt
    from oracle import jt_oracle as o
    for dims in [(64, 512, 64), (8, 1024, 16)]:
        S, R, Cn = dims
        x = o.fill_uniform(2 * S * R * Cn, seed=9)
        want = o.complex_forward_3d(x, S, R, Cn)
        for dev in (0, 1, 0):
            a = x.copy()
            jt.DoubleFFT_3D(S, R, Cn, device=dev).complexForward(a)
            assert o.rel_l2(a, want) < 1e-12 * 24
    y = o.fill_uniform(8192, seed=1)
    for dev in (1, 0):
        b = y.copy()
        jt.DoubleDCT_1D(8192, device=dev).forward(b, True)
        assert o.rel_l2(b, o.dct_forward_nd(y, (8192,), True)) < 1e-12 * 13


@pytest.mark.parametrize("prec", ["Float", "Double"])
def test_batched_bluestein_split_over_devices(prec):
    """BASELINE config 3 behind the drop-in call: a batch of prime-length transforms on ONE host array, split into
    contiguous blocks over the plan's devices (jtb_exec_batch on a multi-GPU plan; no collective).  Virtual ranks on a
    1-GPU box, every GPU otherwise; against the oracle line by line."""
    import numpy as np
    import torch
    import jtransforms_b200 as jt
    from oracle import jt_oracle as o
    n, howmany = 10007, 7                         # prime length -> Bluestein; 7 lines over 2..8 devices: ragged blocks
    ng = torch.cuda.device_count()
    for devices in ([0, 0, 0], list(range(ng)) if ng > 1 else [0, 0]):
        dt = np.float32 if prec == "Float" else np.float64
        x = o.fill_uniform(2 * n * howmany, seed=12, lo=-1.0, hi=1.0).astype(dt)
        a = x.copy()
        f = getattr(jt, prec + "FFT_1D")(n, devices=devices)
        f.complexForwardBatch(a, howmany, 2 * n)
        tol = (1e-5 if prec == "Float" else 1e-12) * 14
        for b in range(howmany):
            want = o.complex_forward_1d(x[2 * n * b:2 * n * (b + 1)].astype(np.float64), n)
            assert o.rel_l2(a[2 * n * b:2 * n * (b + 1)].astype(np.float64), want) < tol, (devices, b)
