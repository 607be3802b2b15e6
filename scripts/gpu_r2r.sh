mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "r2r or dct" 2>&1 | tail -3
for w in dct2d_8192; do
python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1])
print(d['config']['workload'], 'ms/step', d['ms_per_step'], 'GF', round(d['value']), 'roof', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
"; done
JTB_NO_COLPAIR=1 python bench.py --workload dct2d_8192 --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no colpair', d['ms_per_step'])"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,launch__grid_size,launch__block_size,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none -s 11 -c 6 --csv --log-file gpurun_out/k_dct2d_8192.csv python bench.py --workload dct2d_8192 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > /dev/null 2>&1
