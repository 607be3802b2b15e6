mkdir -p gpurun_out
python scripts/dbg_r2r.py 2>&1 | grep " 4096 "
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none -s 14 -c 8 --csv --log-file gpurun_out/k_dct2d_8192.csv python bench.py --workload dct2d_8192 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > /dev/null 2>&1
ncu --metrics $M --clock-control none -s 14 -c 8 --csv --log-file gpurun_out/k_fft2d_real_4096.csv python bench.py --workload fft2d_real_4096 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > /dev/null 2>&1
ncu --metrics $M --clock-control none -s 8 -c 4 --csv --log-file gpurun_out/k_fft1d_2p20.csv python bench.py --workload fft1d_2p20 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --graph off > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "r2r or dct" > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2r.log
