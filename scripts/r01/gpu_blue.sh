#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/blue_ab.log; : > $out
B="bench.py --workload bluestein_f32 --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
for cfg in "" "JTB_BLUE_LOGE=4" "JTB_BLUE_SPLIT=1" "JTB_BLUE_SPLIT=1 JTB_BLUE_LOGE=4"; do
  echo "== $cfg" >> $out
  env $cfg timeout 300 python $B 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['config'], 'ms/step', round(d['ms_per_step'], 3))" >> $out
done
cat $out
