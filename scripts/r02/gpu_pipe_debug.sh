#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_pipe_debug.log
: > $L
NG=$(nvidia-smi -L | wc -l)
run() { echo "== $*" >> $L; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 scripts/pipe_debug_worker.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep "phases\|FAILED" | tail -20 >> $L; }
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=4
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=4 JTB_PIPE_PRIO=0
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=2
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=8
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=1
cat $L
