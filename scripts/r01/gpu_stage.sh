mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "staged or register or batch or offsets" > gpurun_out/pytest_stage.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_stage.log
nproc
echo "== JTB_STAGE=0 (driver pageable path)"; JTB_STAGE=0 timeout 120 python scripts/bench_pageable.py 2>&1 | tee gpurun_out/bench_pageable_off.log
for t in 2 4 8; do echo "== staged, $t threads"; JTB_STAGE_THREADS=$t timeout 120 python scripts/bench_pageable.py 2>&1 | head -1 | tee -a gpurun_out/bench_pageable_on.log; done
