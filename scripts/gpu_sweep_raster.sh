#!/bin/bash
# Sweep of the CTA rasterisation knob on the 512^3 strided axis passes.
mkdir -p gpurun_out
out=gpurun_out/sweep_raster.log
: > $out
for cfg in "" "JTB_RASTER=2" "JTB_RASTER=4" "JTB_RASTER=8" "JTB_RASTER=16" "JTB_RASTER=64" "JTB_RASTER=512" "JTB_RASTER=128" "JTB_RASTER=4096" ; do
  env $cfg REPS=10 timeout 120 python scripts/prof_fft3d.py >> $out 2>&1
done
cat $out
