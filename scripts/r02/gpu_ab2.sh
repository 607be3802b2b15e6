#!/bin/sh
mkdir -p gpurun_out
L=gpurun_out/ab2.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python scripts/ab_strided.py >> $L 2>&1; }
run3() { echo "== fft3d $*" >> $L; env "$@" timeout 300 python scripts/ab_fft3d.py >> $L 2>&1; }
run JTB_TMA=1 JTB_TMA_NG=2 JTB_TMA_NST=3 JTB_TMA_OUT=0
run JTB_TMA=1 JTB_TMA_NG=2 JTB_TMA_NST=3 JTB_TMA_OUT=1
run JTB_TMA=0 JTB_FAST_PREFETCH=74
run JTB_TMA=0 JTB_FAST_PREFETCH=111
run JTB_TMA=0 JTB_FAST_PREFETCH=222
run3 JTB_XPOSE=0
run3 JTB_XPOSE=1
run3 JTB_XPOSE=1 JTB_FAST_PREFETCH=148
run3 JTB_XPOSE=0 JTB_FAST_PREFETCH=148
run3 JTB_XPOSE=1 JTB_FAST_PREFETCH=111
run3 AB_PREC=f32 JTB_XPOSE=0
run3 AB_PREC=f32 JTB_XPOSE=1
run3 AB_PREC=f32 JTB_TMA=1 JTB_TMA_OUT=1
run3 AB_PREC=f32 JTB_TMA=1 JTB_TMA_OUT=0
cat $L
