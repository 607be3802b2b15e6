#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_pipe_debug2.log
: > $L
NG=$(nvidia-smi -L | wc -l)
run() { echo "== $*" >> $L; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 6 --warmup 3 --no-verify 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep "dbg\|Error\|metric" | cut -c1-400 | head -40 >> $L; }
run JTB_BENCH_DEBUG=1 JTB_SLAB_CHUNKS=4
run JTB_BENCH_DEBUG=1 JTB_SLAB_CHUNKS=4 JTB_BENCH_NO_CLOCKS=1
run JTB_SLAB_CHUNKS=4 JTB_BENCH_NO_CLOCKS=1
cat $L
