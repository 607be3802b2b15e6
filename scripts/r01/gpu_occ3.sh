#!/bin/bash
mkdir -p gpurun_out
export KINDS=DCT
for lib in "" "jtransforms_b200/libjtb200_occ3.so"; do
  echo "== lib=$lib"
  if [ -z "$lib" ]; then python scripts/bench_r2r.py 2>&1 | grep "8192, 8192\|4096, 4096\], \"fwd\|256, 256, 256"; else python scripts/with_lib.py $lib scripts/bench_r2r.py 2>&1 | grep "8192, 8192\|256, 256, 256"; fi
done
