#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_ab_rows.log
: > $L
run() { w=$1; shift; env "$@" timeout 300 python scripts/ab_cfg.py $w 2>&1 | tail -1 >> $L; }
run fft1d_2p20 JTB_X=0
run fft2d_real_4096 JTB_X=0
run fft2d_real_4096 JTB_ROW_LOGE=4
run fft2d_real_4096 JTB_ROW_LOGE=3
run dct2d_8192 JTB_X=0
run dct2d_8192 JTB_ROW_LOGE=3
run fft2d_4096_f32 JTB_X=0
cat $L
