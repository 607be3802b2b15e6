mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 30 --warmup 3 --exchange p2p > gpurun_out/bench${n}_p2p.log 2> gpurun_out/bench${n}_p2p.err; echo "bench$n p2p rc=$?"; cat gpurun_out/bench${n}_p2p.log; tail -3 gpurun_out/bench${n}_p2p.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 8 --steps 30 --warmup 3 --exchange nccl > gpurun_out/bench8_nccl.log 2> gpurun_out/bench8_nccl.err; echo "bench8 nccl rc=$?"; cat gpurun_out/bench8_nccl.log
