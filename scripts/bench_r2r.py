"""Device-resident timings of 2-D/3-D DCT/DST/DHT forward and inverse over the size classes (CUDA events)."""
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

cases = [("Double", (256, 256)), ("Double", (512, 512)), ("Double", (1024, 1024)), ("Double", (2048, 2048)), ("Double", (4096, 4096)),
         ("Double", (8192, 8192)), ("Float", (1024, 1024)), ("Float", (4096, 4096)), ("Double", (256, 256, 256))]
kinds = os.environ.get("KINDS", "DCT").split(",")
if os.environ.get("ONLY"):
    cases = [c for c in cases if "x".join(map(str, c[1])) == os.environ["ONLY"] and c[0] == os.environ.get("PREC", "Double")]
for prec, dims in cases:
    n = 1
    for d in dims: n *= d
    dt = torch.float64 if prec == "Double" else torch.float32
    a = torch.rand(n, dtype=dt, device="cuda")
    for kind in kinds:
        plan = getattr(jt, "%s%s_%dD" % (prec, kind, len(dims)))(*dims)
        fwd = (lambda: plan.forward(a)) if kind == "DHT" else (lambda: plan.forward(a, True))
        ms_f = timeit(fwd)
        ms_i = timeit(lambda: plan.inverse(a, True))
        sweep = 2 * n * a.element_size() * len(dims)        # one read + one write of the array per axis
        print(json.dumps({"kind": prec + kind, "dims": dims, "fwd_ms": round(ms_f, 4), "inv_ms": round(ms_i, 4),
                          "fwd_gflops": round(2.5 * n * math.log2(n) / ms_f / 1e6, 1),
                          "fwd_frac_of_axis_sweeps": round(sweep / 6553.9e9 / (ms_f * 1e-3), 3),
                          "inv_frac_of_axis_sweeps": round(sweep / 6553.9e9 / (ms_i * 1e-3), 3)}), flush=True)
    del a
    torch.cuda.empty_cache()
