#!/bin/bash
# N=1 full bench line (other configs, verify, cpu baselines), then the sharded line on all GPUs of the box; argument = tag
TAG=${1:-x}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_bench1_$TAG.json 2> gpurun_out/r02_bench1_$TAG.err
tail -c 400 gpurun_out/r02_bench1_$TAG.err
if [ "$NG" -ge 2 ]; then
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 3 ) > gpurun_out/r02_bench${NG}_$TAG.json 2> gpurun_out/r02_bench${NG}_$TAG.err
  tail -c 600 gpurun_out/r02_bench${NG}_$TAG.err
fi
