#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batch or real or bluestein" > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_batch.log
for mb in 0 256 64; do
  JTB_BATCH_MB=$mb timeout 600 python bench.py --workload bluestein_f32 --steps 3 --warmup 3 --no-cpu --e2e-steps 3 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('JTB_BATCH_MB=$mb', 'ms/step', round(d['ms_per_step'], 2), 'e2e', d['e2e'])"
done
timeout 300 python scripts/bench_real.py 2>&1 | grep "256, 256, 256\|1024, 1024"
