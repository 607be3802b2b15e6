mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 30 --warmup 3 > gpurun_out/bench${n}_p2p.log 2> gpurun_out/bench${n}_p2p.err; echo "bench$n rc=$?"; grep "^{" gpurun_out/bench${n}_p2p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'ms', d['ms_per_step'], 'GF', round(d['value']), 'e2e ms', d['e2e']['ms_per_step'], 'model frac', d['roofline'].get('frac_of_model'), [ (p['pass'][:12], round(p['ms'],4)) for p in d['roofline']['passes']], 'nvlink', d['roofline']['nvlink']['achieved'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 8 --steps 5 --warmup 3 --workload bluestein_f32 > gpurun_out/bench8_blue.log 2> gpurun_out/bench8_blue.err; echo "blue8 rc=$?"; grep "^{" gpurun_out/bench8_blue.log | cut -c1-700; tail -2 gpurun_out/bench8_blue.err | cut -c1-200
