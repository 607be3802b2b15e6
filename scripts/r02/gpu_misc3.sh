#!/bin/bash
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "beyond_2p31" > gpurun_out/r02_pytest_2p31.log 2>&1; tail -3 gpurun_out/r02_pytest_2p31.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $NG --workload bluestein_f32 --steps 5 --warmup 3 > gpurun_out/r02_bench_bluestein_${NG}gpu.json 2> gpurun_out/r02_bench_bluestein_${NG}gpu.err
  tail -c 300 gpurun_out/r02_bench_bluestein_${NG}gpu.err; cut -c1-900 gpurun_out/r02_bench_bluestein_${NG}gpu.json
  timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "bluestein" 2>&1 | tail -2
fi
