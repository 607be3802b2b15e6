#!/bin/bash
bash scripts/gpu_check.sh
timeout 600 python scripts/bench_misc.py > gpurun_out/bench_misc.log 2>&1; cat gpurun_out/bench_misc.log
