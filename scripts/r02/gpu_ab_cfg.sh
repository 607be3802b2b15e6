#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_ab_cfg.log
: > $L
run() { w=$1; shift; env "$@" timeout 300 python scripts/ab_cfg.py $w $VER 2>&1 | tail -1 >> $L; }
VER=v
run fft1d_2p20 JTB_X=0
VER=
for la in 9 10 11; do for w in 4 8; do run fft1d_2p20 JTB_FS_LA=$la JTB_FS_W=$w; done; done
VER=v
run fft2d_real_4096 JTB_X=0
for mb in 8 16 32 64; do run fft2d_real_4096 JTB_STRIP_MB=$mb; done
run dct2d_8192 JTB_X=0
for mb in 8 16 32 64; do run dct2d_8192 JTB_STRIP_MB=$mb; done
VER=
run dht2d_8192 JTB_X=0
run dht2d_8192 JTB_STRIP_MB=16
run fft2d_4096_f32 JTB_X=0
run fft2d_4096_f32 JTB_TMA=0
run fft3d_512_f32 JTB_X=0
cat $L
