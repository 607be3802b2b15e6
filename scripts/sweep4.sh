mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "r2r or dct" > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2r.log
for mb in 8 16 32 64 128 10000; do
echo "strip_mb=$mb"; JTB_STRIP_MB=$mb python bench.py --workload dct2d_8192 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), d['gpu_launches'])"
done
python bench.py --workload fft1d_2p20 --steps 50 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fft1d graph', d['ms_per_step'], round(d['value']), d['gpu_launches'])"
python bench.py --workload fft2d_real_4096 --steps 20 --warmup 3 --no-cpu --graph on 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fft2d real graph', d['ms_per_step'], round(d['value']), d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_dct.csv python bench.py --workload dct2d_8192 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 20 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --workload fft2d_real_4096 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > /dev/null 2>&1
