mkdir -p gpurun_out
timeout 900 python bench.py --workload bluestein_f32 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_bluestein_f32.log 2> gpurun_out/bench_bluestein_f32.err; echo rc=$?; cut -c1-1500 gpurun_out/bench_bluestein_f32.log; tail -3 gpurun_out/bench_bluestein_f32.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q 2>&1 | tail -3
