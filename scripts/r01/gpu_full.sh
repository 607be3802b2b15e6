#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "full or real" > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_full.log
timeout 300 python scripts/bench_real.py 2>&1 | tee gpurun_out/bench_full_new.log
JTB_NO_FASTFULL=1 timeout 300 python scripts/bench_real.py 2>&1 | tee gpurun_out/bench_full_old.log
