for w in dct2d_8192 fft2d_real_4096; do
python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['config']['workload'][:50], 'ms/step', round(d['ms_per_step'],4), 'GF', round(d['value']), 'roof', round(d['roofline']['frac'],3))"
done
python -m pytest tests/test_gpu_parity.py -x -q -k "r2r or dct" 2>&1 | tail -2
