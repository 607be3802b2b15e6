#!/bin/sh
# one gpurun call: lean kernel vs TMA variants vs L2 prefetch on the strided 512^3 passes
mkdir -p gpurun_out
L=gpurun_out/ab_tma.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python scripts/ab_strided.py >> $L 2>&1; }
run JTB_TMA=0
run JTB_TMA=1 JTB_TMA_NG=2 JTB_TMA_NST=3 JTB_TMA_OUT=0
run JTB_TMA=1 JTB_TMA_NG=1 JTB_TMA_NST=3 JTB_TMA_OUT=0
run JTB_TMA=1 JTB_TMA_NG=1 JTB_TMA_NST=3 JTB_TMA_OUT=1
run JTB_TMA=1 JTB_TMA_NG=2 JTB_TMA_NST=3 JTB_TMA_OUT=1
run JTB_TMA=1 JTB_TMA_NG=2 JTB_TMA_NST=2 JTB_TMA_OUT=0
run JTB_TMA=1 JTB_TMA_NG=1 JTB_TMA_NST=2 JTB_TMA_OUT=0
run JTB_TMA=0 JTB_FAST_PREFETCH=296
run JTB_TMA=0 JTB_FAST_PREFETCH=592
run JTB_TMA=0 JTB_FAST_PREFETCH=148
run JTB_TMA=0 AB_PREC=f32
run JTB_TMA=1 AB_PREC=f32 JTB_TMA_OUT=0
run JTB_TMA=1 AB_PREC=f32 JTB_TMA_OUT=1
cat $L
