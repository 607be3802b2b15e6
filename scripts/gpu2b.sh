mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
for ch in 1 2 4 8; do
echo "chunks=$ch"; JTB_SLAB_CHUNKS=$ch timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 30 --warmup 3 --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print(d['ms_per_step'], round(d['value']))"
done
