#!/bin/bash
# full GPU suite + r2r timings + per-kernel launch list of the 8192^2 DCT/DST forward+inverse
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
KINDS=DCT timeout 300 python scripts/bench_r2r.py > gpurun_out/bench_r2r_new2.log 2>&1; cat gpurun_out/bench_r2r_new2.log
ONLY=8192x8192 KINDS=DCT,DST timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/k_r2r_8192.csv python scripts/bench_r2r.py > gpurun_out/ncu_r2r.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/k_r2r_8192.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:70]),{})[r[mi]]=r[vi]
seen={}
for (i,k),m in d.items():
    key=k
    seen.setdefault(key,[]).append((m.get('gpu__time_duration.sum'),m.get('dram__bytes_read.sum'),m.get('dram__bytes_write.sum')))
for k,v in seen.items(): print(k, len(v), v[-1])
PY
