#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
L=gpurun_out/r02_pipe_share_$NG.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29552 scripts/pipe_debug_worker.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep "phases\|FAILED\|rror" | tail -3 >> $L; }
for k in 74 148 222; do
  run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=4 JTB_PIPE_K1_CTAS=$k
  run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=4 JTB_PIPE_K1_CTAS=$k JTB_SCATTER_TMA=1
done
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=8 JTB_PIPE_K1_CTAS=148
run DBG_N=512 DBG_SYNC=0 DBG_STEPS=4 JTB_SLAB_CHUNKS=4 JTB_PIPE_K1_CTAS=148 JTB_PIPE_PRIO=0
cat $L
