mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2_p2p.log 2> gpurun_out/bench2_p2p.err; echo "bench2 rc=$?"; cat gpurun_out/bench2_p2p.log; tail -3 gpurun_out/bench2_p2p.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench2_ref.log 2> gpurun_out/bench2_ref.err; echo "ref rc=$?"; cat gpurun_out/bench2_ref.log | cut -c1-400
