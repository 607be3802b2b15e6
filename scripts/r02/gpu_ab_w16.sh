#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_ab_w16.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python scripts/ab_fft3d.py >> $L 2>&1; }
run JTB_X=0
run JTB_FAST_WS=16
run JTB_FAST_WS=16 JTB_FAST_PREFETCH=55
run JTB_FAST_WS=8 JTB_FAST_PREFETCH=90
run JTB_FAST_WS=8 JTB_FAST_PREFETCH=130
cat $L
