#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "three_pass or fft1d" > gpurun_out/pytest_big.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_big.log
python - <<'PY'
import torch, time, math, sys
sys.path.insert(0, '.')
import jtransforms_b200 as jt
for logn in (21, 22, 24, 26, 27, 28):
    n = 1 << logn
    a = torch.rand(2 * n, dtype=torch.float64, device="cuda")
    f = jt.DoubleFFT_1D(n)
    f.complexForward(a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f.complexForward(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("2^%d: %.3f ms, %.0f GFLOP/s, %.2f sweeps at peak" % (logn, ms, 5 * n * logn / ms / 1e6, ms * 1e-3 * 6553.9e9 / (32 * n)))
    del a, f
    torch.cuda.empty_cache()
PY
