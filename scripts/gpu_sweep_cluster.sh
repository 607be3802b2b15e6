#!/bin/bash
# Sweep of the strided-pass tuning knobs (cluster launch, L2::256B load hint) on the 512^3 axis passes.
mkdir -p gpurun_out
out=gpurun_out/sweep_cluster.log
: > $out
for cfg in "" "JTB_LDHINT=1" "JTB_CLUSTER=2" "JTB_CLUSTER=4" "JTB_CLUSTER=8" "JTB_CLUSTER=2 JTB_LDHINT=1" "JTB_CLUSTER=4 JTB_LDHINT=1" "JTB_FAST_WS=4 JTB_CLUSTER=2" "JTB_FAST_WS=4 JTB_LDHINT=1"; do
  env $cfg REPS=10 timeout 120 python scripts/prof_fft3d.py >> $out 2>&1
done
cat $out
