mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -k "fft3d or slab" 2>&1 | tail -3
for team in 8 16 32; do echo "team=$team"; JTB_TEAM=$team python scripts/prof_fft3d_full.py; done
JTB_NO_SLICE2D=1 python scripts/prof_fft3d_full.py
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_slice.log 2>gpurun_out/bench_slice.err; cut -c1-330 gpurun_out/bench_slice.log; tail -2 gpurun_out/bench_slice.err
