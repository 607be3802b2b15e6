#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 3 ) > gpurun_out/r02_bench${NG}_d.json 2> gpurun_out/r02_bench${NG}_d.err
tail -c 300 gpurun_out/r02_bench${NG}_d.err
cut -c1-1200 gpurun_out/r02_bench${NG}_d.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $NG --steps 2 --warmup 1 2>/dev/null | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02_pytest_dist_$NG.log 2>&1; tail -2 gpurun_out/r02_pytest_dist_$NG.log
