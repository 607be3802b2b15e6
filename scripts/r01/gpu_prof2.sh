mkdir -p gpurun_out
REPS=1 ncu --set full --clock-control none --import-source on -k regex:fft_fast -s 1 -c 5 -o gpurun_out/prof_fast_v1 -f python scripts/prof_fft3d.py > gpurun_out/ncu_fast_v1.log 2>&1
tail -2 gpurun_out/ncu_fast_v1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
for w in fft1d_2p20 fft2d_real_4096 dct2d_8192 bluestein_f32; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; cat gpurun_out/bench_$w.log | cut -c1-900; tail -2 gpurun_out/bench_$w.err
done
