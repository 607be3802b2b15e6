"""A/B of the strided 512-point passes of a 512^3 array (k2: stride C, k1: stride R*C): lean kernel vs the
persistent TMA variants.  Variant = environment (JTB_TMA, JTB_TMA_NG, JTB_TMA_NST, JTB_TMA_OUT, JTB_FAST_PREFETCH);
one process per variant because the library reads the knobs once.  cuFFT (torch.fft) is the checker only."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jtransforms_b200 import _lib

lib = _lib.get()
prec = _lib.F64 if os.environ.get("AB_PREC", "f64") == "f64" else _lib.F32
dt = torch.float64 if prec == _lib.F64 else torch.float32
S = R = Cn = int(os.environ.get("AB_N", "512"))
dev = torch.device("cuda", 0)
a = torch.empty(2 * S * R * Cn, dtype=dt, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def fill():
    _lib.check(lib.jtb_fill_uniform_device(prec, 0, C.c_void_p(a.data_ptr()), a.numel(), 2, -1.0, 1.0, st))


def lines(n, nl, c0, d0, d3, stride):
    _lib.check(lib.jtb_lines_c2c_device(prec, 0, C.c_void_p(a.data_ptr()), n, nl, c0, d0, d3, stride, 0, 1.0, st))


passes = {"k2": (R, Cn * S, Cn, 1, R * Cn, Cn, 1), "k1": (S, R * Cn, R * Cn, 1, S * R * Cn, R * Cn, 0)}
out = {"variant": {k: v for k, v in os.environ.items() if k.startswith("JTB_")}, "n": S, "prec": os.environ.get("AB_PREC", "f64")}
for name, (n, nl, c0, d0, d3, stride, dim) in passes.items():
    fill()
    x = torch.view_as_complex(a.view(S, R, Cn, 2)).clone()
    lines(n, nl, c0, d0, d3, stride)
    y = torch.view_as_complex(a.view(S, R, Cn, 2))
    ref = torch.fft.fft(x, dim=dim)
    err = float((y - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
    del x, ref
    fill()
    for _ in range(3):
        lines(n, nl, c0, d0, d3, stride)
    torch.cuda.synchronize()
    best, tot = 1e9, 0.0
    reps = 10
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lines(n, nl, c0, d0, d3, stride)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
        if _ == 4:
            fill()
    gb = 2 * a.numel() * a.element_size() / 1e9
    out[name] = {"rel_l2": err, "ms_best": best, "ms_avg": tot / reps, "GBps_best": gb / best * 1e3}
print(json.dumps(out))
