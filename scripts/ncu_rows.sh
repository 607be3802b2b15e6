mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"fft_r2r_row" -s 3 -c 1 -o gpurun_out/prof_dctrow -f python bench.py --workload dct2d_8192 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_dctrow.log 2>&1
tail -1 gpurun_out/ncu_dctrow.log
