mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 30 --warmup 3 --e2e-steps 1 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', 'ms', round(d['ms_per_step'],4), 'GF', round(d['value']), [ (p['pass'][:10], round(p['ms'],4)) for p in d['roofline']['passes']], 'nvlink', round(d['roofline']['nvlink']['achieved']))"; }
run 29671 "W8 default"
JTB_SCATTER_W=16 run 29672 "W16"
JTB_SLICE2D=1 run 29673 "slice2d fused"
JTB_SLICE2D=1 JTB_TEAM=8 run 29674 "slice2d team8"
