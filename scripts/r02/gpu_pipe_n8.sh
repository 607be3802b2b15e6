#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
L=gpurun_out/r02_pipe_ab_$NG.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $NG --steps 20 --warmup 3 --e2e-steps 1 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep "JtbError\|metric" | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln)
        print(json.dumps({k: d[k] for k in ('ms_per_step', 'value', 'verified', 'rel_l2')}), json.dumps(d['roofline']['passes']))
    else:
        print(ln[:300])
" >> $L; }
run JTB_X=0
run JTB_SLAB_CHUNKS=4 JTB_SCATTER_TMA=1
run JTB_SLAB_CHUNKS=8 JTB_SCATTER_TMA=1
run JTB_SLAB_CHUNKS=4 JTB_SCATTER_TMA=1 JTB_PIPE_K1_CTAS=74
run JTB_SLAB_CHUNKS=4
cat $L
