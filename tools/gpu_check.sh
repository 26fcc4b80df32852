#!/bin/bash
# usage: tools/gpu_check.sh <tag> [pytest args]   -- runs the gpu tests, the C3 bench and a launch list on the GPU box
TAG=$1; shift
/usr/local/graft/bin/gpurun --timeout 900 -- "mkdir -p gpurun_out; timeout 600 python -m pytest tests -m gpu -x -q $* > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-200; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 26 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1" 2>&1 | tail -5
python - <<PY
import csv
rows=[r for r in csv.reader(open('/root/repo/gpurun_out/launches_$TAG.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg={}
for r in rows[1:27]:
    k=r[ki].split('(')[0].replace('b200::','')
    agg.setdefault(k,[]).append(float(r[vi].replace(',',''))/1e6)
tot=0
for k,v in agg.items():
    per=sum(v)/max(1,len(v)) if not ('pyramid' in k or 'halfpyr' in k) else sum(v)/ (len(v)/ (7 if 'k_pyramid'==k else 4))
    tot+=per
    print('%-18s %.3f ms'%(k,per))
print('sum %.3f ms'%tot)
PY
