#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_b.log
echo "== two-pass mixed radix" > gpurun_out/r02_bench_misc.log
MISC_1D_ONLY=1 timeout 300 python scripts/bench_misc.py >> gpurun_out/r02_bench_misc.log 2>&1
echo "== Bluestein route (JTB_NO_MIXED2=1)" >> gpurun_out/r02_bench_misc.log
MISC_1D_ONLY=1 JTB_NO_MIXED2=1 timeout 300 python scripts/bench_misc.py >> gpurun_out/r02_bench_misc.log 2>&1
cat gpurun_out/r02_bench_misc.log
L=gpurun_out/r02_ab_dht.log
: > $L
run() { w=$1; shift; env "$@" timeout 300 python scripts/ab_cfg.py $w v 2>&1 | tail -1 >> $L; }
run dht2d_8192 JTB_X=0
run dht2d_8192 JTB_NO_DHTFOLD=1
cat $L
