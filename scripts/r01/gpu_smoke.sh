mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fftw_golden or r2r_fused or staged or three_pass_against" 2>&1 | tail -2
