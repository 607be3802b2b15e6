mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -s 11 -c 4 -o gpurun_out/prof_cfg2 -f python bench.py --workload fft2d_real_4096 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > gpurun_out/ncu_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fft_fast2|fft_conv" -s 12 -c 3 -o gpurun_out/prof_cfg3 -f python bench.py --workload bluestein_f32 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fft_fast2" -s 6 -c 2 -o gpurun_out/prof_cfg1 -f python bench.py --workload fft1d_2p20 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > gpurun_out/ncu_cfg1.log 2>&1
python bench.py --workload bluestein_f32 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bluestein', d['ms_per_step'], round(d['value']), d['gpu_launches'])"
