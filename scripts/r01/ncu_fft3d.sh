mkdir -p gpurun_out
REPS=1 ncu --set full --clock-control none --import-source on -k regex:fft_tile -s 3 -c 3 -o gpurun_out/prof_v1 -f python scripts/prof_fft3d.py > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_v1.log
