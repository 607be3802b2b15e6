mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fft2d_complex or fft3d_complex" > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_t.log
