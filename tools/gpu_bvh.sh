#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/bvh_build_bench.py 1000000 16000000 2>&1 | tee gpurun_out/bvh_build.log
echo "=== ncu launch list of the 1M build"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bvh1m.csv python tools/bvh_build_bench.py 1000000 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_bvh1m.csv')) if r and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-50:]; val=float(r[-1].replace(',',''))/1e3
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=val
for k,v in agg.items(): print('%-52s n=%3d total %9.1f us  per launch %8.1f us'%(k,v[0],v[1],v[1]/v[0]))
PY
