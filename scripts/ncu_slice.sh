mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fft_slice2d -s 1 -c 1 -o gpurun_out/prof_slice2d -f python scripts/prof_fft3d_full.py > gpurun_out/ncu_slice.log 2>&1
tail -2 gpurun_out/ncu_slice.log
