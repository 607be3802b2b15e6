#!/bin/bash
# 8-GPU validation: sharded bench line (device-resident + single-process e2e over all GPUs), reference arm, dist tests
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r02_topo_$NG.txt 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 3 ) > gpurun_out/r02_bench${NG}_c.json 2> gpurun_out/r02_bench${NG}_c.err
tail -c 300 gpurun_out/r02_bench${NG}_c.err
cut -c1-1500 gpurun_out/r02_bench${NG}_c.json
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02_pytest_dist_$NG.log 2>&1; tail -3 gpurun_out/r02_pytest_dist_$NG.log
