#!/bin/bash
# usage: tools/gpu_check.sh <tag> [kernel regex for a full ncu capture]   -- gpu tests, C3 bench, launch list (+ ncu --set full) on the GPU box
TAG=$1; KREGEX=$2
FULL=""
if [ -n "$KREGEX" ]; then FULL="timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:$KREGEX' -s 4 -c 3 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1;"; fi
/usr/local/graft/bin/gpurun --timeout 1200 -- "mkdir -p gpurun_out; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-200; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 94 -c 29 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; $FULL" 2>&1 | tail -5
python - <<PY
import csv, json
rows=[r for r in csv.reader(open('/root/repo/gpurun_out/launches_$TAG.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg={}
for r in rows[1:30]:
    k=r[ki].split('(')[0].replace('b200::','')
    agg.setdefault(k,[]).append(float(r[vi].replace(',',''))/1e6)
tot=0
for k,v in agg.items():
    per=sum(v)/max(1,len(v)) if not ('pyramid' in k or 'halfpyr' in k) else sum(v)/ (len(v)/ (7 if 'k_pyramid'==k else 4))
    tot+=per
    print('%-18s %.3f ms'%(k,per))
print('sum %.3f ms'%tot)
try:
    d=json.loads(open('/root/repo/gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
    print('value %.0f fps  %.3f ms/step   e2e %.0f fps   stages %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v,3) for k,v in d['roofline']['extractor_stage_ms'].items()}))
except Exception as e:
    print('bench line unreadable', e)
PY
