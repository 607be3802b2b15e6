mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "register or batch" > gpurun_out/pytest_pin.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pin.log
timeout 120 python scripts/bench_pageable.py 2>&1 | tee gpurun_out/bench_pageable.log
