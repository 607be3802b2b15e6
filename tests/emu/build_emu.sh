#!/bin/sh
# Builds libjtb200_emu.so: the library's real sources compiled by g++ against the CUDA-semantics shim
# (tests/emu/emu_cuda.h).  Test infrastructure only -- never loaded by the product package.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../jtransforms_b200/csrc
OUT=$HERE/_build
mkdir -p "$OUT"
rm -f "$OUT"/*.o
FLAGS="-O2 -g -std=c++17 -fPIC -DJTB_EMU_BUILD -I$HERE -I$SRC -x c++ -pthread -Wno-unknown-pragmas"
for f in jtb_ctx tile_f64 tile_f32 jtb_capi jtb_fast jtb_fast2 jtb_mixed jtb_r2r_inv jtb_stage jtb_tma jtb_slab; do
  g++ $FLAGS -c "$SRC/$f.cu" -o "$OUT/$f.o" &
done
g++ -O2 -g -std=c++17 -fPIC -I$HERE -pthread -c "$HERE/emu_cuda.cpp" -o "$OUT/emu_cuda.o" &
wait
for f in jtb_ctx tile_f64 tile_f32 jtb_capi jtb_fast jtb_fast2 jtb_mixed jtb_r2r_inv jtb_stage jtb_tma jtb_slab emu_cuda; do test -f "$OUT/$f.o" || { echo "error: $f failed"; exit 1; }; done
g++ -shared -pthread -o "$OUT/libjtb200_emu.so" "$OUT"/jtb_ctx.o "$OUT"/tile_f64.o "$OUT"/tile_f32.o "$OUT"/jtb_capi.o "$OUT"/jtb_fast.o "$OUT"/jtb_fast2.o "$OUT"/jtb_mixed.o "$OUT"/jtb_r2r_inv.o "$OUT"/jtb_stage.o "$OUT"/jtb_tma.o "$OUT"/jtb_slab.o "$OUT"/emu_cuda.o -ldl
echo built "$OUT/libjtb200_emu.so"
