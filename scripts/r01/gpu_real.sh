#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_real.py > gpurun_out/bench_real_new.log 2>&1; cat gpurun_out/bench_real_new.log
JTB_NO_RFFTINV=1 timeout 300 python scripts/bench_real.py > gpurun_out/bench_real_old.log 2>&1; cat gpurun_out/bench_real_old.log
