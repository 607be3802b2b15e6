mkdir -p gpurun_out
for pipe in 0 1 0 1; do
JTB_SLICE2D_PIPE=$pipe timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$pipe bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('PIPE=$pipe ms/step', round(d['ms_per_step'], 4))"
done
