#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_ab_blue.log
: > $L
for sp in 1 2 3 0; do
  for mb in 1024 256; do
    echo "== JTB_BLUE_SPLIT=$sp JTB_BLUE_MB=$mb" >> $L
    JTB_BLUE_SPLIT=$sp JTB_BLUE_MB=$mb timeout 300 python - >> $L 2>&1 <<PY
import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import jtransforms_b200 as jt
from oracle import jt_oracle as o
n, b = 1000003, 512
plan = jt.FloatFFT_1D(n)
x = o.fill_uniform(2 * n * 2, seed=7, lo=-1.0, hi=1.0).astype(np.float32)
t = torch.from_numpy(x.copy()).cuda()
plan.complexForwardBatch(t, 2, 2 * n)
got = t.cpu().numpy().astype(np.float64)
want = np.concatenate([o.complex_forward_1d(x[i * 2 * n:(i + 1) * 2 * n].astype(np.float64), n) for i in range(2)])
err = o.rel_l2(got, want)
a = torch.rand(2 * n * b, dtype=torch.float32, device="cuda")
for _ in range(2): plan.complexForwardBatch(a, b, 2 * n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): plan.complexForwardBatch(a, b, 2 * n)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"us_per_transform": ms * 1e3 / b, "ms_4096": ms / b * 4096, "rel_l2": err}))
PY
  done
done
cat $L
