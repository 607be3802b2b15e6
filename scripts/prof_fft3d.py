"""Times the three axis passes of the 512^3 transform separately (CUDA events), for tuning sweeps."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jtransforms_b200 import _lib
from jtransforms_b200.dist import SlabFFT3D
S = R = Cn = int(os.environ.get("N3D", "512"))
slab = SlabFFT3D(S, R, Cn)
a = torch.rand(2 * S * R * Cn, dtype=torch.float64, device="cuda")
passes = [("k3", (Cn, S * R, 1, 0, Cn, 1)), ("k2", (R, Cn * S, Cn, 1, R * Cn, Cn)), ("k1", (S, R * Cn, R * Cn, 1, S * R * Cn, R * Cn))]
out = {}
reps = int(os.environ.get("REPS", "5"))
for name, g in passes:
    slab._lines(a, *g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        slab._lines(a, *g)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name] = {"ms": round(ms, 4), "GBps": round(2 * a.numel() * 8 / ms / 1e6, 1)}
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("JTB_")}, "passes": out}))
