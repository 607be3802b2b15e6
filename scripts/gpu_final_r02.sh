#!/bin/bash
# final validation of the round: smoke, full GPU suite, default bench line (+ reference arm), on whatever the box has
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_bench1_final.json 2> gpurun_out/r02_bench1_final.err; tail -c 200 gpurun_out/r02_bench1_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_benchref1_final.json 2>/dev/null; cut -c1-300 gpurun_out/r02_benchref1_final.json
if [ "$NG" -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/r02_bench${NG}_final.json 2> gpurun_out/r02_bench${NG}_final.err
  cut -c1-400 gpurun_out/r02_bench${NG}_final.json
fi
