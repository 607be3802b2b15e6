#!/bin/bash
# ncu --set full of the forward and inverse DCT row kernels (8192^2)
mkdir -p gpurun_out
export ONLY=8192x8192 KINDS=DCT
ncu --set full --clock-control none --import-source on -k regex:fft_r2r_row_kernel -s 3 -c 1 -o gpurun_out/prof_dctrow_v2 -f python scripts/bench_r2r.py > gpurun_out/ncu_dctrow_v2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_r2r_row_inv_kernel -s 3 -c 1 -o gpurun_out/prof_dctrowinv_v2 -f python scripts/bench_r2r.py > gpurun_out/ncu_dctrowinv_v2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
