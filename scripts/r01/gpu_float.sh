#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_float.py 2>&1 | tee gpurun_out/bench_float.log
PRECS=Float REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/k_float.csv python scripts/bench_float.py > gpurun_out/ncu_float.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/k_float.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
order=[]
for r in rows[1:]:
    k=(int(r[ii]), r[ki].split('(')[0][-62:])
    if k not in d: d[k]={}; order.append(k)
    d[k][r[mi]]=r[vi]
# print the last 3 launches' worth per distinct kernel: keep last occurrence of each kernel name
last={}
for k in order: last[k[1]]=(k[0], d[k])
for name,(i,m) in sorted(last.items(), key=lambda x:x[1][0]):
    print(i, name, int(m['gpu__time_duration.sum'])//1000, 'us', int(m['dram__bytes_read.sum'])>>20, '+', int(m['dram__bytes_write.sum'])>>20, 'MiB')
PY
