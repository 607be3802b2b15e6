"""Device-resident timings of non-config sizes (reference benchmark defaults, fft/BenchmarkDoubleFFT.java:58-62)."""
import os, sys, json, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtransforms_b200 as jt

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

cases = [("1d", (10368,)), ("1d", (27000,)), ("1d", (75600,)), ("1d", (165375,)), ("1d", (362880,)), ("1d", (1562500,)), ("1d", (3211264,)),
         ("1d", (6250000,)), ("1d", (1 << 22,)), ("1d", (1 << 24,)),
         ("2d", (1050, 1050)), ("2d", (1960, 1960)), ("2d", (1024, 1024)), ("2d", (8192, 8192)),
         ("3d", (95, 95, 95)), ("3d", (180, 180, 180)), ("3d", (420, 420, 420)), ("3d", (256, 256, 256)), ("3d", (1024, 1024, 1024))]
if os.environ.get("MISC_1D_ONLY"):
    cases = [c for c in cases if c[0] == "1d"]
for kind, dims in cases:
    n = 1
    for d in dims: n *= d
    a = torch.rand(2 * n, dtype=torch.float64, device="cuda")
    plan = {"1d": jt.DoubleFFT_1D, "2d": jt.DoubleFFT_2D, "3d": jt.DoubleFFT_3D}[kind](*dims)
    t0 = time.perf_counter(); plan.complexForward(a); torch.cuda.synchronize(); first = time.perf_counter() - t0
    ms = timeit(lambda: plan.complexForward(a), reps=3 if n > 1e8 else 10)
    print(json.dumps({"dims": dims, "ms": round(ms, 4), "gflops": round(5 * n * math.log2(n) / ms / 1e6, 1), "GBps_1sweep": round(32 * n / ms / 1e6, 1), "first_call_s": round(first, 3)}))
    del a, plan
    torch.cuda.empty_cache()
