"""Whole 512^3 (or AB_N^3) complexForward through jtb_exec_device: ms per transform + parity vs cuFFT (checker only).
Variant = environment knobs (JTB_XPOSE, JTB_FAST_PREFETCH, JTB_TMA ...); one process per variant."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jtransforms_b200 as jt
from jtransforms_b200 import _lib

lib = _lib.get()
f64 = os.environ.get("AB_PREC", "f64") == "f64"
prec = _lib.F64 if f64 else _lib.F32
dt = torch.float64 if f64 else torch.float32
N = int(os.environ.get("AB_N", "512"))
dev = torch.device("cuda", 0)
a = torch.empty(2 * N ** 3, dtype=dt, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
plan = (jt.DoubleFFT_3D if f64 else jt.FloatFFT_3D)(N, N, N, device=0)


def fill():
    _lib.check(lib.jtb_fill_uniform_device(prec, 0, C.c_void_p(a.data_ptr()), a.numel(), 2, -1.0, 1.0, st))


fill()
x = torch.view_as_complex(a.view(N, N, N, 2)).clone()
plan.complexForward(a)
ref = torch.fft.fftn(x)
y = torch.view_as_complex(a.view(N, N, N, 2))
err = float((y - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
del x, ref
fill()
for _ in range(3):
    plan.complexForward(a)
torch.cuda.synchronize()
fill()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    plan.complexForward(a)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"variant": {k: v for k, v in os.environ.items() if k.startswith("JTB_") or k.startswith("AB_")},
                  "rel_l2": err, "ms": e0.elapsed_time(e1) / reps}))
