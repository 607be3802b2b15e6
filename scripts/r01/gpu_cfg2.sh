mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in fft1d_2p20 fft2d_real_4096 dct2d_8192 bluestein_f32; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1])
print(d['config']['workload'], 'ms/step', d['ms_per_step'], 'GF', round(d['value']), 'roof', round(d['roofline']['frac'],3), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/bench_$w.err
done
for mb in 16 32 128 512; do echo "blue_mb=$mb"; JTB_BLUE_MB=$mb python bench.py --workload bluestein_f32 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), d['gpu_launches'])"; done
