#!/bin/bash
# pipelined exchange A/B on all GPUs of the box: JTB_SLAB_CHUNKS = 1 (fused slice kernel + barrier + k1) vs 2/4/8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
L=gpurun_out/r02_pipe_ab_$NG.log
: > $L
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02_pytest_dist_$NG.log 2>&1; tail -3 gpurun_out/r02_pytest_dist_$NG.log
for ch in 1 2 4 8; do
  echo "== JTB_SLAB_CHUNKS=$ch" >> $L
  JTB_SLAB_CHUNKS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2953$ch bench.py --gpus $NG --steps 20 --warmup 3 --e2e-steps 2 2>> gpurun_out/r02_pipe_ab_$NG.err | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln)
        print(json.dumps({k: d[k] for k in ('ms_per_step', 'value', 'verified', 'rel_l2')}), json.dumps(d['roofline']['passes']), json.dumps(d['e2e']))
" >> $L
done
cat $L
