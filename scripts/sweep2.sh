mkdir -p gpurun_out
( for ws in 8 4; do for wc in 8 4 2 1; do JTB_FAST_WS=$ws JTB_FAST_WC=$wc python scripts/prof_fft3d.py; done; done ) > gpurun_out/sweep2.log 2>&1
cat gpurun_out/sweep2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fft3d or fft2d_complex or fft1d_batch" 2>&1 | tail -3
