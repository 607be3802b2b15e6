mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_dist.log
