mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for w in fft3d_512 fft1d_2p20 fft2d_real_4096 dct2d_8192 bluestein_f32; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1])
print(d['config']['workload'][:50], 'ms/step', round(d['ms_per_step'],4), 'GF', round(d['value']), 'roof', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/bench_$w.err
done
