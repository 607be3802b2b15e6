#!/bin/bash
# parity of the rewritten row kernels + A/B of the row radix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "r2r or dct2d or real" > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2r.log
out=gpurun_out/rows_ab2.log
: > $out
run() { echo "== $1" >> $out; shift; "$@" 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'fwd_ms' in d: print(d['kind'], d['dims'], 'fwd', d['fwd_ms'], 'inv', d['inv_ms'])
    else: print(d['config']['workload'], d['ms_per_step'])
" >> $out; }
B2="bench.py --workload fft2d_real_4096 --steps 20 --warmup 3 --no-cpu --e2e-steps 1"
export KINDS=DCT
run "default r2r" python scripts/bench_r2r.py
run "loge3 r2r" env JTB_ROW_LOGE=3 python scripts/bench_r2r.py
run "loge4 r2r" env JTB_ROW_LOGE=4 python scripts/bench_r2r.py
run "default rfft2d" python $B2
run "loge4 rfft2d" env JTB_ROW_LOGE=4 python $B2
cat $out
