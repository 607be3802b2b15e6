mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_dist.log
for ex in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex > gpurun_out/bench2_$ex.log 2> gpurun_out/bench2_$ex.err; echo "bench2 $ex rc=$?"; cat gpurun_out/bench2_$ex.log; tail -5 gpurun_out/bench2_$ex.err
done
