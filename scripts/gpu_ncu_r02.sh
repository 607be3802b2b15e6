#!/bin/bash
mkdir -p gpurun_out
# (1) launch list of the bench command (per-launch durations, cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fft3d_512.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-verify --e2e-steps 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/r02_bench_under_ncu.log
# (2) full captures: the three passes of the double transform (second repetition), the float passes (TMA kernel)
REPS=2 ncu --set full --clock-control none --import-source on -k regex:fft_fast_kernel -s 3 -c 3 -o gpurun_out/r02_prof_fft3d_f64 -f python scripts/prof_plan3d.py > gpurun_out/r02_ncu_f64.log 2>&1; tail -2 gpurun_out/r02_ncu_f64.log
PROF_PREC=f32 REPS=2 ncu --set full --clock-control none --import-source on -k regex:"fft_tma_kernel|fft_fast_kernel" -s 3 -c 3 -o gpurun_out/r02_prof_fft3d_f32 -f python scripts/prof_plan3d.py > gpurun_out/r02_ncu_f32.log 2>&1; tail -2 gpurun_out/r02_ncu_f32.log
# (3) mixed two-pass after the 1024-thread variant
MISC_1D_ONLY=1 timeout 300 python scripts/bench_misc.py > gpurun_out/r02_bench_misc_big.log 2>&1; cat gpurun_out/r02_bench_misc_big.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "smooth or golden or nonpow2 or fft1d" > gpurun_out/r02_pytest_mixed_big.log 2>&1; tail -2 gpurun_out/r02_pytest_mixed_big.log
ls -la gpurun_out/*.ncu-rep | tail -3
