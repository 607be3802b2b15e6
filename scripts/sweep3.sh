mkdir -p gpurun_out
for mb in 8 16 24 32 48 64 1000; do
echo "strip_mb=$mb"; JTB_STRIP_MB=$mb python bench.py --workload fft2d_real_4096 --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']))"
done
python bench.py --workload fft1d_2p20 --steps 50 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fft1d graph', d['ms_per_step'], round(d['value']), d['gpu_launches'])"
python bench.py --workload fft1d_2p20 --steps 50 --warmup 3 --no-cpu --graph off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fft1d eager', d['ms_per_step'], round(d['value']), d['gpu_launches'])"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fft2d" 2>&1 | tail -2
