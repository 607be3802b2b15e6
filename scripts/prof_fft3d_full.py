"""Times DoubleFFT_3D.complexForward 512^3 (device resident) and its in-slice part separately."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jtransforms_b200 import _lib
from jtransforms_b200.dist import SlabFFT3D
S = R = Cn = 512
slab = SlabFFT3D(S, R, Cn)
lib = _lib.get()
a = torch.rand(2 * S * R * Cn, dtype=torch.float64, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
t2d = timeit(lambda: _lib.check(lib.jtb_fft2d_slices_device(0, 0, C.c_void_p(a.data_ptr()), S, R, Cn, 1, 0, None, 0, st)))
a.uniform_()
tall = timeit(lambda: slab.forward(a))
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("JTB_")}, "slices2d_ms": round(t2d, 4), "full_ms": round(tall, 4)}))
