#!/usr/bin/env python
"""bench.py -- headline benchmark of jtransforms_b200 (contract: see the task statement / DESIGN.md section 6).

Default workload = BASELINE.json's target configuration: DoubleFFT_3D.complexForward on 512^3 complex
doubles (2 GiB, in place).  One "step" = one transform.  With N ranks the SAME transform is slab-decomposed
over the N GPUs (scaling = "strong") with the all-to-all fused into the k2 pass over NVLink.

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus 8 --steps 20 --warmup 3
  python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU arm)

What the JSON line holds (all measured in this run):
  value / ms_per_step  device-resident transforms, CUDA events on the launching stream, max over ranks
  e2e                  the drop-in call -- DoubleFFT_3D(...).complexForward(host array) = jtb_exec -- on ONE pinned host
                       array in natural order, H2D + kernels + D2H inside the timed region.  With N > 1 rank 0 drives all
                       N GPUs from one process (jtb_plan_set_devices: slab g over GPU g's PCIe link) while the other
                       ranks wait: that is the call a JTransforms user makes.
  verified / rel_l2    one fresh step of the timed path (and of the e2e path) compared with the oracle, element for
                       element, tolerance 1e-12 * log2(N)
  roofline             per-pass CUDA-event timing of the kernels of the step
  cpu_baseline         oracle/jt_ref.c (the JTransforms algorithm restated in C, all host threads) = `value`;
                       pocketfft (scipy.fft) reported beside it
  other_configs        (N = 1 only) BASELINE.json configs 1-4 in the same terms: ms, GFLOP/s, fraction of the sweep
                       model, verified against the oracle, bounded CPU baseline

--workload {fft1d_2p20, fft2d_real_4096, bluestein_f32, dct2d_8192, dst2d_8192, dht2d_8192, fft3d_512} times one of
the other configurations as the main line instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FFT GFLOP/s (5N*log2N/t)"
UNIT = "GFLOP/s"
TOL_F64, TOL_F32 = 1e-12, 1e-5      # north star: relative L2 <= tol * log2(N)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------- workloads
class Workload:
    """name, total points, flops per transform (5 N log2 N complex / 2.5 N log2 N real), element counts"""

    def __init__(self, name, world=1):
        self.name = name
        self.batch = 1
        if name == "fft3d_512":
            self.dims, self.prec, self.kind, self.op = (512, 512, 512), "f64", "fft", "complexForward"
            self.N = 512 ** 3
            self.flops = 5.0 * self.N * math.log2(self.N)
            self.elems = 2 * self.N
            self.desc = "DoubleFFT_3D.complexForward 512^3 (2 GiB, in place)"
            self.sweeps = 3
        elif name == "fft1d_2p20":
            self.dims, self.prec, self.kind, self.op = (1 << 20,), "f64", "fft", "complexForward"
            self.N = 1 << 20
            self.flops = 5.0 * self.N * 20
            self.elems = 2 * self.N
            self.desc = "DoubleFFT_1D.complexForward n=2^20"
            self.sweeps = 2
        elif name == "fft2d_real_4096":
            self.dims, self.prec, self.kind, self.op = (4096, 4096), "f64", "fft", "realForward"
            self.N = 4096 * 4096
            self.flops = 2.5 * self.N * 24
            self.elems = self.N
            self.desc = "DoubleFFT_2D.realForward 4096x4096"
            self.sweeps = 2
        elif name in ("dct2d_8192", "dst2d_8192", "dht2d_8192"):
            self.kind = name[:3]
            self.dims, self.prec, self.op = (8192, 8192), "f64", "forward"
            self.N = 8192 * 8192
            self.flops = 2.5 * self.N * 26
            self.elems = self.N
            self.desc = "Double%s_2D.forward%s 8192x8192" % (self.kind.upper(), "" if self.kind == "dht" else "(scale=true)")
            self.sweeps = 2
        elif name == "bluestein_f32":
            self.dims, self.prec, self.kind, self.op = (1000003,), "f32", "fft", "complexForward"
            self.total_batch = 4096                      # BASELINE.json config 3; split evenly over the ranks
            self.batch = self.total_batch // world        # transforms per GPU (32.8 GB / world of input in HBM)
            self.e2e_batch = 256                          # host-buffer leg: 2 GB per rank (rate metric, bounded)
            self.N = 1000003
            self.flops = 5.0 * self.N * math.log2(self.N) * self.batch
            self.elems = 2 * self.N * self.batch
            self.desc = "FloatFFT_1D.complexForward n=1000003 (Bluestein), batch 4096 split over the GPUs"
            self.sweeps = 4
        else:
            raise SystemExit("unknown workload " + name)
        self.esize = 8 if self.prec == "f64" else 4
        self.bytes = self.elems * self.esize
        self.tol = (TOL_F64 if self.prec == "f64" else TOL_F32) * math.log2(self.N)

    def model_bytes(self):
        """algorithmic HBM bytes of one step: `sweeps` read+write passes over the working set"""
        if self.name == "bluestein_f32":
            return (16.0 * self.N + 56.0 * (1 << 21)) * self.batch   # (8n+8M) + 16M + (16M+8M) + (8M+8n), M = 2^21
        return 2.0 * self.sweeps * self.bytes


# ------------------------------------------------------------------------------- clocks sampler
class Clocks(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, False, []

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------- CPU arm
def _best_of(f, budget, max_reps=5):
    f()
    b, used, reps = 1e30, 0.0, 0
    while used < budget and reps < max_reps:
        t0 = time.perf_counter()
        f()
        dt = time.perf_counter() - t0
        b = min(b, dt)
        used += dt
        reps += 1
    return b


def cpu_reference(w: Workload, budget_s: float = 20.0):
    """The reference's algorithm on the host cores, bounded sample.  `value` is ALWAYS the restatement of the JTransforms
    algorithm (oracle/jt_ref.c: Ooura-style radix-4 passes + the reference's thread partitioning; kind = "port" -- the
    reference itself is Java and cannot run here); scipy.fft (pocketfft, a different and usually faster CPU library) is
    reported beside it under "pocketfft" so that both ratios can be read off."""
    import numpy as np
    import scipy.fft as sfft
    from oracle import cref
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(2)
    if w.name == "fft3d_512":
        x = rng.random(2 * 512 ** 3)
        port = (lambda: cref.cfft3d(x, 512, 512, 512, -1, cores), w.flops,
                "full 512^3 transform, oracle/jt_ref.c (JTransforms algorithm restated in C), %d threads" % cores, cores)
        z = x.view(np.complex128).reshape(512, 512, 512)
        alt = (lambda: sfft.fftn(z, workers=cores, overwrite_x=True), w.flops,
               "full 512^3 transform, scipy.fft.fftn (pocketfft) workers=%d" % cores)
    elif w.name == "fft1d_2p20":
        x = rng.random(2 << 20)
        nt = min(4, cores)          # the reference never uses more than 4 threads for 1-D (CommonUtils.java:3727-3734)
        port = (lambda: cref.cfft1d(x, 1 << 20, -1, nt), w.flops,
                "full size, oracle/jt_ref.c, %d threads (reference maximum for 1-D)" % nt, nt)
        z1 = x.view(np.complex128)
        alt = (lambda: sfft.fft(z1, workers=nt), w.flops, "full size, scipy.fft.fft (pocketfft) workers=%d" % nt)
    elif w.name == "fft2d_real_4096":
        x = rng.random(4096 * 4096)
        port = (lambda: cref.rfft2d(x, 4096, 4096, cores), w.flops,
                "full size, oracle/jt_ref.c realForward 2-D (row rdft + column cdft + rdft2d_sub), %d threads" % cores, cores)
        xr = x.reshape(4096, 4096)
        alt = (lambda: sfft.rfft2(xr, workers=cores), w.flops, "full size, scipy.fft.rfft2 workers=%d" % cores)
    elif w.name in ("dct2d_8192", "dst2d_8192", "dht2d_8192"):
        rows = 2048                                                # a quarter of the rows, full-length 8192-point lines
        x = rng.random(rows * 8192)
        n = float(x.size)
        fl = 2.5 * n * math.log2(n)
        port = (lambda: cref.r2r2d(x, rows, 8192, w.kind, cores), fl,
                "%dx8192 (a quarter of the rows), oracle/jt_ref.c %s 2-D (Makhoul FFT + gather-columns driver), %d threads"
                % (rows, w.kind.upper(), cores), cores)
        xr = x.reshape(rows, 8192)
        if w.kind == "dht":
            alt = (lambda: sfft.fft2(xr, workers=cores), fl, "%dx8192, scipy.fft.fft2 (a DHT costs one real FFT) workers=%d" % (rows, cores))
        else:
            ty = 2
            f = sfft.dctn if w.kind == "dct" else sfft.dstn
            alt = (lambda: f(xr, type=ty, norm="ortho", workers=cores), fl, "%dx8192, scipy.fft.%sn workers=%d" % (rows, w.kind, cores))
    else:
        nb = 4
        x = rng.random(nb * 2 * 1000003).astype(np.float32)
        fl = 5.0 * 1000003 * math.log2(1000003) * nb
        port = (lambda: cref.bluestein_f32(x, 1000003, nb, cores), fl,
                "%d transforms of n=1000003 complex64, oracle/jt_ref.c Bluestein (per-call ak buffer, M = 2^21), %d threads"
                % (nb, cores), cores)
        z = x.view(np.complex64).reshape(nb, 1000003)
        alt = (lambda: sfft.fft(z, axis=-1, workers=cores), fl, "%d transforms, scipy.fft.fft workers=%d" % (nb, cores))
    bp = _best_of(port[0], budget_s / 2)
    ba = _best_of(alt[0], budget_s / 2)
    return {"value": port[1] / bp / 1e9, "unit": UNIT, "cores": port[3], "kind": "port", "sample": port[2],
            "seconds": bp,
            "pocketfft": {"value": alt[1] / ba / 1e9, "unit": UNIT, "sample": alt[2], "seconds": ba}}


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference(w, budget_s=30.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": w.prec, "data": "synthetic",
            "config": {"workload": w.desc, "note": "CPU arm: " + base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "pocketfft")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- helpers (GPU arm)
def oracle_of(w: Workload, x):
    """expected output of one step on input x (numpy, FP64), from the oracle"""
    from oracle import jt_oracle as o
    if w.name == "fft3d_512":
        return o.complex_forward_3d(x, *w.dims)
    if w.name == "fft1d_2p20":
        return o.complex_forward_1d(x, w.dims[0])
    if w.name == "fft2d_real_4096":
        return o.real_forward_2d(x, *w.dims)
    if w.name == "dct2d_8192":
        return o.dct_forward_nd(x, w.dims, True)
    if w.name == "dst2d_8192":
        return o.dst_forward_nd(x, w.dims, True)
    if w.name == "dht2d_8192":
        return o.dht_forward_nd(x, w.dims)
    raise ValueError(w.name)


def make_step(jt, w: Workload, a, device):
    klass = {"fft3d_512": jt.DoubleFFT_3D, "fft1d_2p20": jt.DoubleFFT_1D, "fft2d_real_4096": jt.DoubleFFT_2D,
             "dct2d_8192": jt.DoubleDCT_2D, "dst2d_8192": jt.DoubleDST_2D, "dht2d_8192": jt.DoubleDHT_2D,
             "bluestein_f32": jt.FloatFFT_1D}[w.name]
    plan = klass(*w.dims, device=device)
    if w.name == "bluestein_f32":
        return plan, (lambda t=a: plan.complexForwardBatch(t, w.batch, 2 * w.N))
    if w.name in ("dct2d_8192", "dst2d_8192"):
        return plan, (lambda t=a: plan.forward(t, True))
    if w.name == "dht2d_8192":
        return plan, (lambda t=a: plan.forward(t))
    if w.name == "fft2d_real_4096":
        return plan, (lambda t=a: plan.realForward(t))
    return plan, (lambda t=a: plan.complexForward(t))


def verify_single(torch, jt, w: Workload, step_on, dev):
    """one fresh step on seeded data against the oracle (element for element)"""
    import numpy as np
    from oracle import jt_oracle as o
    if w.name == "bluestein_f32":
        # two lines of the batch against the FP64 DFT of the same float inputs
        nb = 2
        x = o.fill_uniform(nb * 2 * w.N, seed=7, lo=-1.0, hi=1.0).astype(np.float32)
        t = torch.from_numpy(x.copy()).to(dev)
        plan = jt.FloatFFT_1D(w.N, device=dev.index)
        plan.complexForwardBatch(t, nb, 2 * w.N)
        got = t.cpu().numpy().astype(np.float64)
        want = np.concatenate([o.complex_forward_1d(x[i * 2 * w.N:(i + 1) * 2 * w.N].astype(np.float64), w.N) for i in range(nb)])
        err = o.rel_l2(got, want)
        return bool(err <= w.tol), float(err)
    x = o.fill_uniform(w.elems, seed=2, lo=-1.0, hi=1.0)
    t = torch.from_numpy(x).to(dev)
    step_on(t)
    torch.cuda.synchronize()
    err = o.rel_l2(t.cpu().numpy(), oracle_of(w, x))
    return bool(err <= w.tol), float(err)


def time_steps(torch, step, steps, warmup, refill=None, refill_every=40):
    for _ in range(warmup):
        step()
    if refill:
        refill()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if refill and i and i % refill_every == 0:
            refill()
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def other_config(torch, jt, lib, _lib, name, dev, hbm_peak, with_cpu, steps=10, warmup=3):
    """one of BASELINE.json's configs 1-4 on this GPU: device-resident ms, GFLOP/s, fraction of the sweep model,
    oracle check, bounded CPU baseline"""
    w = Workload(name)
    tdt = torch.float64 if w.prec == "f64" else torch.float32
    prec = _lib.F64 if w.prec == "f64" else _lib.F32
    a = torch.empty(w.elems, dtype=tdt, device=dev)

    def fill():
        _lib.check(lib.jtb_fill_uniform_device(prec, dev.index, C.c_void_p(a.data_ptr()), a.numel(), 2, -1.0, 1.0,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    fill()
    plan, step = make_step(jt, w, a, dev.index)
    l0 = lib.jtb_launch_count(dev.index)
    step()
    per_step = lib.jtb_launch_count(dev.index) - l0
    ms = time_steps(torch, step, steps, warmup, fill, refill_every=4 if w.prec == "f32" or w.kind != "fft" else 40)
    out = {"workload": w.desc, "ms_per_step": ms, "value": w.flops / (ms * 1e-3) / 1e9, "unit": UNIT, "dtype": w.prec,
           "steps": steps, "warmup": warmup, "gpu_launches_per_step": int(per_step)}
    if name == "fft1d_2p20":
        # launch-bound: the same step replayed from a CUDA graph
        s = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(s):
            step()
            torch.cuda.synchronize()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, stream=s):
                step()
            msg = time_steps(torch, cg.replay, steps * 4, warmup, None)
        out["ms_per_step_cuda_graph"] = msg
        out["value_cuda_graph"] = w.flops / (msg * 1e-3) / 1e9
    algo = w.model_bytes()
    best = min(ms, out.get("ms_per_step_cuda_graph", ms))
    out["roofline"] = {"bound": "hbm", "model": "%d-sweep model (SURVEY.md 8(d))" % w.sweeps, "algorithmic_bytes": algo,
                       "achieved": algo / (best * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": algo / (best * 1e-3) / 1e9 / hbm_peak}
    if 2 * w.bytes < 126e6:
        out["l2"] = "working set %.0f MiB fits the 126 MB L2 (as in the reference's repeated-call benchmark)" % (w.bytes / 2 ** 20)
    del a
    ok, err = verify_single(torch, jt, w, (lambda t: make_step(jt, w, t, dev.index)[1]()), dev)
    out["verified"], out["rel_l2"], out["tolerance"] = ok, err, w.tol
    if with_cpu:
        cpu = cpu_reference(w, budget_s=6.0)
        out["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "pocketfft")}
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="jtb200")
    ap.add_argument("--workload", default="fft3d_512")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs array")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle checks")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--exchange", default="auto", help="3-D slab exchange: p2p (fused peer stores) | nccl")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w = Workload(args.workload, world if args.impl != "reference" else 1)
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import jtransforms_b200 as jt
    from jtransforms_b200 import _lib
    from jtransforms_b200.dist import SlabFFT3D
    from oracle import jt_oracle as o          # the checker (verified / rel_l2), never on the timed path

    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.get()
    hbm_peak, peak_kind = peaks()
    tdt = torch.float64 if w.prec == "f64" else torch.float32
    prec = _lib.F64 if w.prec == "f64" else _lib.F32
    cur_stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fill(t, seed):
        _lib.check(lib.jtb_fill_uniform_device(prec, local, C.c_void_p(t.data_ptr()), t.numel(), seed, -1.0, 1.0, cur_stream()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sharded = w.name == "fft3d_512" and world > 1
    if world > 1 and not sharded and w.name != "bluestein_f32":
        raise SystemExit("workload %s does not shard: replicas only (run with --gpus 1)" % w.name)

    # ---- the step (device resident)
    slab = None
    if sharded:
        S, R, Cn = w.dims
        slab = SlabFFT3D(S, R, Cn, prec, device_index=local, exchange=args.exchange)
        a = torch.empty(slab.local_elements(), dtype=tdt, device=dev)
        step = lambda: slab.forward(a)
    else:
        a = torch.empty(w.elems, dtype=tdt, device=dev)
        plan, step = make_step(jt, w, a, local)
    local_bytes = a.numel() * w.esize
    seed0 = 2 + rank * a.numel()
    fill(a, seed0)
    refill_every = 4 if (w.prec == "f32" or w.kind != "fft") else 40   # repeated in-place forward transforms grow

    dbg = os.environ.get("JTB_BENCH_DEBUG")

    def dbg_status(where):
        if dbg and slab is not None:
            torch.cuda.synchronize()
            try:
                slab.status()
                print("[dbg] rank %d %s ok" % (rank, where), file=sys.stderr, flush=True)
            except Exception as e:
                print("[dbg] rank %d %s FAILED %s" % (rank, where, e), file=sys.stderr, flush=True)
    for i in range(args.warmup):
        step()
        dbg_status("warmup %d" % i)
    fill(a, seed0)
    barrier()
    dbg_status("after barrier 1")
    clocks = Clocks(local)
    if rank == 0 and os.environ.get("JTB_BENCH_NO_CLOCKS") is None:
        clocks.start()
    l0 = lib.jtb_launch_count(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        if i and i % refill_every == 0:
            fill(a, seed0)
        step()
        if dbg:
            dbg_status("timed %d" % i)
    e1.record()
    barrier()
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = lib.jtb_launch_count(local) - l0
    if slab is not None:
        slab.status()
    total_flops = w.flops * (world if w.name == "bluestein_f32" else 1)
    value = total_flops / (ms_per_step * 1e-3) / 1e9

    # ---- per-pass roofline of the step's kernels, CUDA events on the launching stream
    def timed(fn, reps=5, sync_ranks=False):
        fn()
        barrier() if sync_ranks else torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        barrier() if sync_ranks else torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1) / reps
        return max_over_ranks(t) if sync_ranks else t

    roof = None
    if w.name == "fft3d_512" and world == 1:
        # the three passes exactly as nd_c2c issues them: rows in place, then the two axis-swapping strided passes
        # (k2: a -> work laid out [r][s][c]; k1: work -> a in natural order); JTB_XPOSE=0: both strided passes in place
        S, R, Cn = w.dims
        xpose = os.environ.get("JTB_XPOSE", "1") != "0"
        wk = torch.empty_like(a)
        ap_, wp_ = C.c_void_p(a.data_ptr()), C.c_void_p(wk.data_ptr())

        def lines(ptr, n, nl, c0, d0, d3, st):
            _lib.check(lib.jtb_lines_c2c_device(prec, local, ptr, n, nl, c0, d0, d3, st, 0, 1.0, cur_stream()))

        def lines_out(src, dst, n, nl, c0, d3, st, od3, ost):
            _lib.check(lib.jtb_lines_c2c_out_device(prec, local, src, dst, n, nl, c0, d3, st, od3, ost, 0, 1.0, cur_stream()))
        passes = [("k3 rows (contiguous, in place)", "fft_fast_kernel<double,9,3,contiguous>", lambda: lines(ap_, Cn, S * R, 1, 0, Cn, 1))]
        if xpose:
            passes += [("k2 columns (read stride C, stored axis-swapped [r][s][c])", "fft_fast_kernel<double,9,3,strided,8>",
                        lambda: lines_out(ap_, wp_, R, Cn * S, Cn, R * Cn, Cn, Cn, S * Cn)),
                       ("k1 slices (read stride C from the swapped array, stored in natural order)", "fft_fast_kernel<double,9,3,strided,8>",
                        lambda: lines_out(wp_, ap_, S, Cn * R, Cn, S * Cn, Cn, Cn, R * Cn))]
        else:
            passes += [("k2 columns (stride C, in place)", "fft_fast_kernel<double,9,3,strided,8>", lambda: lines(ap_, R, Cn * S, Cn, 1, R * Cn, Cn)),
                       ("k1 slices (stride R*C, in place)", "fft_fast_kernel<double,9,3,strided,8>", lambda: lines(ap_, S, R * Cn, R * Cn, 1, S * R * Cn, R * Cn))]
        # timed IN SEQUENCE (k3, k2, k1, k3, ... exactly like consecutive steps) with an event between the launches, so
        # that every pass sees the cache / TLB / clock state it has inside the step
        for _, _, fn in passes:
            fn()
        fill(a, seed0)
        reps = max(3, min(args.steps, 10))
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(passes) + 1)] for _ in range(reps)]
        torch.cuda.synchronize()
        for r in range(reps):
            evs[r][0].record()
            for i, (_, _, fn) in enumerate(passes):
                fn()
                evs[r][i + 1].record()
        torch.cuda.synchronize()
        per = []
        for i, (name, kern, _) in enumerate(passes):
            t_ms = sum(evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(reps)) / reps
            per.append({"pass": name, "kernel": kern, "ms": t_ms, "GBps": 2 * local_bytes / (t_ms * 1e-3) / 1e9,
                        "frac": 2 * local_bytes / (t_ms * 1e-3) / 1e9 / hbm_peak})
        del wk
        worst = max(per, key=lambda p: p["ms"])
        roof = {"bound": "hbm", "kernel": worst["kernel"] + " -- " + worst["pass"], "achieved": worst["GBps"],
                "peak": hbm_peak, "peak_source": peak_kind, "unit": "GB/s", "frac": worst["GBps"] / hbm_peak,
                "traffic": 4.238e9, "traffic_source": "profiles/r01_ncu_fft_fast_512_summary.json (ncu --set full, dram read+write per launch)",
                "algorithmic_bytes_per_launch": 2 * local_bytes, "passes": per,
                "whole_step": {"algorithmic_bytes": 6 * local_bytes, "GBps": 6 * local_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": 6 * local_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak}}
    elif sharded:
        # phase timing of the real step on every rank (events inside the library): in-slice passes incl. the exchange
        # stores, waiting for the peers, slice-axis pass; max over ranks
        S, R, Cn = w.dims
        _lib.check(lib.jtb_slab_profile(slab._m, 1))
        acc = [0.0, 0.0, 0.0]
        reps = 5
        for _ in range(reps):
            barrier()
            slab.forward(a)
            torch.cuda.synchronize()
            t3 = (C.c_float * 3)()
            _lib.check(lib.jtb_slab_last_times(slab._m, t3))
            for i in range(3):
                acc[i] += t3[i] / reps
        _lib.check(lib.jtb_slab_profile(slab._m, 0))
        acc = [max_over_ranks(v) for v in acc]
        nv_bytes = (world - 1) / world * local_bytes            # sent per GPU over NVLink
        nv_peak = 770.0                                          # measured peer-copy GB/s per direction (B200_PROFILING.md)
        per = [{"pass": "k3 rows + k2 columns with the all-to-all in the k2 stores (peer stores over NVLink)", "ms": acc[0],
                "hbm_GBps": 4 * local_bytes / (acc[0] * 1e-3) / 1e9, "nvlink_GBps": nv_bytes / (acc[0] * 1e-3) / 1e9},
               {"pass": "wait for the peers (flag barrier)", "ms": acc[1]},
               {"pass": "k1 slices on the re-slabbed block", "ms": acc[2], "GBps": 2 * local_bytes / (acc[2] * 1e-3) / 1e9}]
        model_ms = (6 * local_bytes / (hbm_peak * 1e9) + nv_bytes / 900e9) * 1e3
        overlap_ms = max(6 * local_bytes / (hbm_peak * 1e9), nv_bytes / 900e9) * 1e3
        roof = {"bound": "hbm", "kernel": "fft_fast_kernel<double,9,3,strided,8> -- k1 slices on the re-slabbed block",
                "achieved": per[2]["GBps"], "peak": hbm_peak, "peak_source": peak_kind, "unit": "GB/s",
                "frac": per[2]["GBps"] / hbm_peak, "traffic": None, "algorithmic_bytes_per_launch": 2 * local_bytes,
                "passes": per,
                "nvlink": {"kernel": "fft_slice2d_kernel / fft_scatter_kernel (k2 pass fused with the all-to-all)",
                           "achieved": nv_bytes / ((acc[0] + acc[1]) * 1e-3) / 1e9, "peak": nv_peak,
                           "peak_source": "measured peer copy, B200_PROFILING.md (900 nominal)", "unit": "GB/s",
                           "frac": nv_bytes / ((acc[0] + acc[1]) * 1e-3) / 1e9 / nv_peak, "bytes_sent_per_gpu": nv_bytes},
                "model_ms": model_ms, "frac_of_model": model_ms / ms_per_step, "overlap_bound_ms": overlap_ms,
                "model": "BASELINE.md section 2: 6D/P / HBM peak + (P-1)/P * D/P / 900 GB/s, no overlap"}
    else:
        algo = w.model_bytes()
        gbps = algo / (ms_per_step * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "whole step (%d-sweep model)" % w.sweeps, "achieved": gbps, "peak": hbm_peak,
                "peak_source": peak_kind, "unit": "GB/s", "frac": gbps / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_launch": algo}

    # ---- verification of the timed path: one fresh step on seeded data against the oracle
    verified, rel = None, None
    xh = want = None
    if not args.no_verify:
        if w.name == "fft3d_512":
            S, R, Cn = w.dims
            # rank 0 makes input and expected output; the other ranks receive both over NCCL and check their own block
            xg = torch.empty(w.elems, dtype=tdt, device=dev)
            wg = torch.empty(w.elems, dtype=tdt, device=dev)
            if rank == 0:
                xh = o.fill_uniform(w.elems, seed=2, lo=-1.0, hi=1.0)
                want = o.complex_forward_3d(xh, S, R, Cn)
                xg.copy_(torch.from_numpy(xh))
                wg.copy_(torch.from_numpy(want))
            if world > 1:
                dist.broadcast(xg, 0)
                dist.broadcast(wg, 0)
                a.copy_(xg[rank * a.numel():(rank + 1) * a.numel()])
                res = slab.forward(a)
                Rh = R // world
                mine = wg.view(S, R, 2 * Cn)[:, rank * Rh:(rank + 1) * Rh, :]
                num = float(torch.sum((res.view(S, Rh, 2 * Cn) - mine) ** 2).item())
                den = float(torch.sum(mine ** 2).item())
                t2 = torch.tensor([num, den], dtype=torch.float64, device=dev)
                dist.all_reduce(t2)
                rel = math.sqrt(float(t2[0].item()) / float(t2[1].item()))
                torch.cuda.synchronize()
                slab.status()
            else:
                a.copy_(xg)
                step()
                rel = float((torch.linalg.norm(a - wg) / torch.linalg.norm(wg)).item())
            verified = bool(rel <= w.tol)
            del xg, wg
        elif rank == 0:
            verified, rel = verify_single(torch, jt, w, (lambda t: make_step(jt, w, t, local)[1]()), dev)
    torch.cuda.empty_cache()

    # ---- end to end: ONE pinned host array in natural order through the public API (jtb_exec), H2D + kernels + D2H
    e2e = None
    esteps = max(1, min(args.e2e_steps, args.steps))
    if w.name == "fft3d_512":
        S, R, Cn = w.dims
        barrier()
        if rank == 0:
            hp = C.c_void_p()
            _lib.check(lib.jtb_host_alloc(C.byref(hp), w.bytes))
            harr = np.ctypeslib.as_array((C.c_double * w.elems).from_address(hp.value))
            f3 = jt.DoubleFFT_3D(S, R, Cn, devices=list(range(world))) if world > 1 else jt.DoubleFFT_3D(S, R, Cn, device=local)
            if xh is not None:
                harr[:] = xh
            else:
                harr[:] = 0.5
            f3.complexForward(harr)
            e2e_ok, e2e_rel = None, None
            if want is not None:
                e2e_rel = float(o.rel_l2(harr, want))
                e2e_ok = bool(e2e_rel <= w.tol)
            harr[:] = 0.25
            f3.complexForward(harr)
            harr[:] = 0.25
            t0 = time.perf_counter()
            for _ in range(esteps):
                f3.complexForward(harr)
            dt = (time.perf_counter() - t0) / esteps
            del f3
            lib.jtb_host_free(hp)
            e2e = {"value": w.flops / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": w.bytes, "d2h_bytes_per_step": w.bytes,
                   "ms_per_step": dt * 1e3, "steps": esteps, "verified": e2e_ok, "rel_l2": e2e_rel,
                   "api": "DoubleFFT_3D(512,512,512%s).complexForward(pinned host array) = jtb_exec; one process drives %d GPU(s), "
                          "slab g over GPU g's PCIe link, result in natural [S][R][C] order" % (", devices=0..%d" % (world - 1) if world > 1 else "", world),
                   # H2D and D2H are serial (every output depends on every input): each of the `world` PCIe links moves
                   # bytes/world per direction in (step - kernels)/2
                   "pcie_GBps_per_link_per_direction": (w.bytes / world) / (max(dt - ms_per_step * 1e-3, 1e-9) / 2) / 1e9}
        barrier()
    else:
        hp = C.c_void_p()
        e_elems = w.elems if w.name != "bluestein_f32" else 2 * w.N * w.e2e_batch
        e_bytes = e_elems * w.esize
        _lib.check(lib.jtb_host_alloc(C.byref(hp), e_bytes))
        ct = C.c_double if w.prec == "f64" else C.c_float
        harr = np.ctypeslib.as_array((ct * e_elems).from_address(hp.value))
        harr[:] = 0.25
        if w.name == "bluestein_f32":
            e2e_step = lambda: plan.complexForwardBatch(harr, w.e2e_batch, 2 * w.N)
        else:
            e2e_step = make_step(jt, w, harr, local)[1]
        e2e_step()
        harr[:] = 0.25
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            e2e_step()
        dt = max_over_ranks((time.perf_counter() - t0) / esteps)
        lib.jtb_host_free(hp)
        e_flops = total_flops if w.name != "bluestein_f32" else 5.0 * w.N * math.log2(w.N) * w.e2e_batch * world
        e2e = {"value": e_flops / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": e_bytes,
               "d2h_bytes_per_step": e_bytes, "ms_per_step": dt * 1e3, "steps": esteps}
        if w.name == "bluestein_f32":
            e2e["note"] = "host-buffer leg on %d transforms per rank (2 GB); rate metric" % w.e2e_batch

    clk = None
    if rank == 0:
        clocks.stop_flag = True
        if clocks.is_alive():
            clocks.join(timeout=2)
        clk = clocks.summary()

    # ---- the other BASELINE.json configurations (1 GPU only: they do not shard)
    others = None
    if rank == 0 and world == 1 and not args.no_others and w.name == "fft3d_512":
        del a
        torch.cuda.empty_cache()
        others = []
        for name in ("fft1d_2p20", "fft2d_real_4096", "bluestein_f32", "dct2d_8192", "dst2d_8192", "dht2d_8192"):
            try:
                others.append(other_config(torch, jt, lib, _lib, name, dev, hbm_peak, not args.no_cpu))
            except Exception as e:      # never lose the main line to a secondary measurement
                others.append({"workload": name, "error": repr(e)})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(w)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "pocketfft")}

    if rank == 0:
        l2note = ("working set %.0f MiB per GPU >> 126 MB L2, no flush needed" % (local_bytes / 2 ** 20)) if local_bytes > 300e6 else \
                 ("working set %.0f MiB per GPU: inputs + outputs of consecutive steps exceed the 126 MB L2 only partly; "
                  "no flush (the reference's benchmark repeats calls on one array the same way)" % (local_bytes / 2 ** 20))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak" if w.name == "bluestein_f32" else "strong", "vs_baseline": None, "dtype": w.prec,
                "data": "synthetic",
                "config": {"workload": w.desc, "l2": l2note,
                           "parallelism": ("slab%d/%s (one process per GPU, CUDA IPC peers)" % (world, slab.exchange)) if sharded else
                                          ("batch/%d" % world if world > 1 else "single")},
                "verified": verified, "rel_l2": rel, "tolerance": w.tol,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        if others is not None:
            line["other_configs"] = others
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
