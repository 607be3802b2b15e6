#!/bin/bash
# parity of the fused inverse / single-pass column kernels + timings before (JTB_NO_FASTINV=1) and after
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "r2r or dct2d" > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2r.log
KINDS=DCT,DST,DHT timeout 300 python scripts/bench_r2r.py > gpurun_out/bench_r2r_new.log 2>&1; cat gpurun_out/bench_r2r_new.log
JTB_NO_FASTINV=1 timeout 300 python scripts/bench_r2r.py > gpurun_out/bench_r2r_old.log 2>&1; cat gpurun_out/bench_r2r_old.log
