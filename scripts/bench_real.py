"""Device-resident timings of realForward / realInverse (2-D, 3-D, batched 1-D) -- CUDA events."""
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

for prec, dims in [("Double", (1024, 1024)), ("Double", (4096, 4096)), ("Double", (8192, 8192)), ("Float", (4096, 4096)),
                   ("Double", (256, 256, 256)), ("Double", (512, 512, 512))]:
    n = 1
    for d in dims: n *= d
    a = torch.rand(n, dtype=torch.float64 if prec == "Double" else torch.float32, device="cuda")
    plan = getattr(jt, "%sFFT_%dD" % (prec, len(dims)))(*dims)
    f = timeit(lambda: plan.realForward(a))
    i = timeit(lambda: plan.realInverse(a, True))
    full = None
    if n <= (1 << 27):
        a2 = torch.rand(2 * n, dtype=a.dtype, device="cuda")
        full = round(timeit(lambda: plan.realForwardFull(a2)), 4)
        del a2
    sweep = 2 * n * a.element_size()
    print(json.dumps({"kind": prec + "FFT real", "dims": dims, "fwd_ms": round(f, 4), "inv_ms": round(i, 4), "full_ms": full,
                      "fwd_gflops": round(2.5 * n * math.log2(n) / f / 1e6, 1),
                      "fwd_sweeps_at_peak": round(f * 1e-3 * 6553.9e9 / sweep, 2), "inv_sweeps_at_peak": round(i * 1e-3 * 6553.9e9 / sweep, 2)}), flush=True)
    del a
    torch.cuda.empty_cache()
