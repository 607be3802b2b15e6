"""Device-resident timings of the Float classes (complex / real, 2-D and 3-D) -- CUDA events."""
import os, sys, json, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtransforms_b200 as jt

def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

reps = int(os.environ.get("REPS", "10"))
for prec in os.environ.get("PRECS", "Float,Double").split(","):
    for dims in [(4096, 4096), (512, 512, 512)]:
        n = 1
        for d in dims: n *= d
        dt = torch.float64 if prec == "Double" else torch.float32
        a = torch.rand(2 * n, dtype=dt, device="cuda")
        plan = getattr(jt, "%sFFT_%dD" % (prec, len(dims)))(*dims)
        c = timeit(lambda: plan.complexForward(a), reps)
        r = timeit(lambda: plan.realForward(a), reps)
        es = a.element_size()
        print(json.dumps({"kind": prec, "dims": dims, "c2c_ms": round(c, 4), "r2c_ms": round(r, 4),
                          "c2c_sweeps_at_peak": round(c * 1e-3 * 6553.9e9 / (4 * n * es), 2),
                          "r2c_sweeps_at_peak": round(r * 1e-3 * 6553.9e9 / (2 * n * es), 2)}), flush=True)
        del a
        torch.cuda.empty_cache()
