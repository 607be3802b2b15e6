mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fft_fast2 -s 9 -c 1 -o gpurun_out/prof_rfftrow -f python bench.py --workload fft2d_real_4096 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > gpurun_out/ncu_rfftrow.log 2>&1
tail -1 gpurun_out/ncu_rfftrow.log
