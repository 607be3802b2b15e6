#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_ab_rowprefetch.log
: > $L
run() { w=$1; shift; env "$@" timeout 300 python scripts/ab_cfg.py $w $VER 2>&1 | tail -1 >> $L; }
VER=
for pf in 0 74 148 296 592; do run dct2d_8192 JTB_ROW_PREFETCH=$pf; done
for pf in 0 148 296; do run fft2d_real_4096 JTB_ROW_PREFETCH=$pf; done
for pf in 148 296; do run dht2d_8192 JTB_ROW_PREFETCH=$pf; done
VER=v
run dct2d_8192 JTB_ROW_PREFETCH=296
cat $L
# sanitizer sweep (memcheck) over the small parity cases, all kernel families incl. this round's
timeout 1500 compute-sanitizer --print-limit 3 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(golden or fft1d_real or fft2d_complex or r2r or smooth) and not 8192 and not 4096 and not 2048 and not 100000 and not 65536 and not 32768 and not 6250000 and not 3211264 and not 1562500" > gpurun_out/r02_sanitizer.log 2>&1
echo "sanitizer rc=$?"; tail -3 gpurun_out/r02_sanitizer.log
