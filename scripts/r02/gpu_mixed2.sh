#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "smooth_reference_sizes or fft1d" > gpurun_out/r02_pytest_mixed2.log 2>&1; tail -3 gpurun_out/r02_pytest_mixed2.log
echo "== two-pass mixed radix" > gpurun_out/r02_bench_misc.log
MISC_1D_ONLY=1 timeout 300 python scripts/bench_misc.py >> gpurun_out/r02_bench_misc.log 2>&1
echo "== Bluestein route (JTB_NO_MIXED2=1)" >> gpurun_out/r02_bench_misc.log
MISC_1D_ONLY=1 JTB_NO_MIXED2=1 timeout 300 python scripts/bench_misc.py >> gpurun_out/r02_bench_misc.log 2>&1
cat gpurun_out/r02_bench_misc.log
# memcheck over the small-size parity cases (every kernel family once)
timeout 1500 compute-sanitizer --print-limit 3 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or fft1d_real or fft2d_complex or r2r and not 8192 and not 4096 and not 2048 and not 100000 and not 65536 and not 32768" > gpurun_out/r02_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r02_sanitizer.log; tail -4 gpurun_out/r02_sanitizer.log
