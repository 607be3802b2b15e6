#!/bin/bash
mkdir -p gpurun_out
JTB_SLICE2D=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -k "fft3d or 512 or slices or scatter" > gpurun_out/pytest_slice.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_slice.log
out=gpurun_out/slice2d_ab.log; : > $out
for cfg in "JTB_SLICE2D=0" "JTB_SLICE2D=1 JTB_SLICE2D_PIPE=0" "JTB_SLICE2D=1 JTB_SLICE2D_PIPE=1" "JTB_SLICE2D=1 JTB_SLICE2D_PIPE=1 JTB_TEAM=32" "JTB_SLICE2D=1 JTB_SLICE2D_PIPE=1 JTB_TEAM=8" "JTB_SLICE2D=1 JTB_SLICE2D_PIPE=1 JTB_TEAM=64"; do
  echo "== $cfg" >> $out
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'], 4), 'launches', d['gpu_launches'])" >> $out
done
cat $out
