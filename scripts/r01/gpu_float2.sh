#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fft2d or fft3d or offsets" > gpurun_out/pytest_nd.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_nd.log
timeout 300 python scripts/bench_float.py 2>&1 | tee gpurun_out/bench_float2.log
echo "== JTB_FAST_WS=8 (float strided 64-byte groups)"
PRECS=Float JTB_FAST_WS=8 timeout 300 python scripts/bench_float.py 2>&1 | tee -a gpurun_out/bench_float2.log
