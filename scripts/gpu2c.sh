mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2_p2p.log 2> gpurun_out/bench2_p2p.err; echo "rc=$?"; grep "^{" gpurun_out/bench2_p2p.log | cut -c1-3000; tail -3 gpurun_out/bench2_p2p.err | cut -c1-300
