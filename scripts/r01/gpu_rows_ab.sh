#!/bin/bash
# A/B of the row-kernel options: radix-16 rows (JTB_ROW_LOGE=4) and derived twiddles (libjtb200_twd.so)
mkdir -p gpurun_out
out=gpurun_out/rows_ab.log
: > $out
TWD=jtransforms_b200/libjtb200_twd.so
run() { echo "== $1" >> $out; shift; "$@" 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'passes' in d: print({k: v['ms'] for k, v in d['passes'].items()})
    elif 'fwd_ms' in d: print(d['kind'], d['dims'], 'fwd', d['fwd_ms'], 'inv', d['inv_ms'])
    else: print(d['config']['workload'], d['ms_per_step'])
" >> $out; }
B2="bench.py --workload fft2d_real_4096 --steps 20 --warmup 3 --no-cpu --e2e-steps 1"
export ONLY=8192x8192 KINDS=DCT,DST
run "default r2r" python scripts/bench_r2r.py
run "loge4 r2r" env JTB_ROW_LOGE=4 python scripts/bench_r2r.py
run "twd r2r" python scripts/with_lib.py $TWD scripts/bench_r2r.py
run "twd+loge4 r2r" env JTB_ROW_LOGE=4 python scripts/with_lib.py $TWD scripts/bench_r2r.py
run "default rfft2d" python $B2
run "loge4 rfft2d" env JTB_ROW_LOGE=4 python $B2
run "twd rfft2d" python scripts/with_lib.py $TWD $B2
run "twd+loge4 rfft2d" env JTB_ROW_LOGE=4 python scripts/with_lib.py $TWD $B2
run "default fft3d" env REPS=10 python scripts/prof_fft3d.py
run "twd fft3d" env REPS=10 python scripts/with_lib.py $TWD scripts/prof_fft3d.py
cat $out
