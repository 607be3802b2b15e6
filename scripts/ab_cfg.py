"""A/B helper: device-resident time of one BASELINE config under the environment's tuning knobs.
usage: python scripts/ab_cfg.py {fft1d_2p20|fft2d_real_4096|dct2d_8192|dst2d_8192|dht2d_8192|fft3d_512|fft3d_512_f32|fft2d_4096_f32} [verify]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jtransforms_b200 as jt
from oracle import jt_oracle as o

name = sys.argv[1]
verify = len(sys.argv) > 2
dev = torch.device("cuda", 0)
f32 = name.endswith("_f32")
dt = torch.float32 if f32 else torch.float64
P = "Float" if f32 else "Double"
if name == "fft1d_2p20":
    n = 1 << 20; elems = 2 * n; plan = jt.DoubleFFT_1D(n); step = lambda t: plan.complexForward(t); want = lambda x: o.complex_forward_1d(x, n)
elif name == "fft2d_real_4096":
    elems = 4096 * 4096; plan = jt.DoubleFFT_2D(4096, 4096); step = lambda t: plan.realForward(t); want = lambda x: o.real_forward_2d(x, 4096, 4096)
elif name in ("dct2d_8192", "dst2d_8192", "dht2d_8192"):
    k = name[:3].upper(); elems = 8192 * 8192; plan = getattr(jt, "Double%s_2D" % k)(8192, 8192)
    step = (lambda t: plan.forward(t)) if k == "DHT" else (lambda t: plan.forward(t, True))
    want = {"DCT": lambda x: o.dct_forward_nd(x, (8192, 8192), True), "DST": lambda x: o.dst_forward_nd(x, (8192, 8192), True),
            "DHT": lambda x: o.dht_forward_nd(x, (8192, 8192))}[k]
elif name.startswith("fft3d_512"):
    elems = 2 * 512 ** 3; plan = getattr(jt, P + "FFT_3D")(512, 512, 512); step = lambda t: plan.complexForward(t)
    want = lambda x: o.complex_forward_3d(x, 512, 512, 512)
elif name.startswith("fft2d_4096"):
    elems = 2 * 4096 ** 2; plan = getattr(jt, P + "FFT_2D")(4096, 4096); step = lambda t: plan.complexForward(t)
    want = lambda x: o.complex_forward_2d(x, 4096, 4096)
else:
    raise SystemExit("unknown " + name)
rel = None
if verify:
    x = o.fill_uniform(elems, seed=2, lo=-1.0, hi=1.0)
    t = torch.from_numpy(x).to(dev).to(dt)
    step(t)
    torch.cuda.synchronize()
    rel = float(o.rel_l2(t.cpu().numpy().astype(np.float64), want(x)))
a = torch.rand(elems, dtype=dt, device=dev)
a0 = a.clone()
for _ in range(3):
    step(a)
best, tot, reps = 1e9, 0.0, 5
for r in range(reps):
    a.copy_(a0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        step(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    best = min(best, ms); tot += ms / reps
knobs = {k: v for k, v in os.environ.items() if k.startswith("JTB_")}
print(json.dumps({"workload": name, "knobs": knobs, "ms_best": round(best, 5), "ms_avg": round(tot, 5), "rel_l2": rel}))
