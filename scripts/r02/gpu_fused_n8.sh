#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
L=gpurun_out/r02_fused_pipe_$NG.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $NG --steps 20 --warmup 3 --e2e-steps 1 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep "JtbError\|metric" | python -c "
import sys, json
seen = set()
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln)
        print(json.dumps({k: d[k] for k in ('ms_per_step', 'value', 'verified', 'rel_l2')}), json.dumps([round(p['ms'], 4) for p in d['roofline']['passes']]), d['e2e']['verified'])
    elif ln[:60] not in seen:
        seen.add(ln[:60]); print(ln[:200])
" >> $L; }
run JTB_SLAB_FUSED=4
run JTB_SLAB_FUSED=8
run JTB_SLAB_FUSED=2
cat $L
