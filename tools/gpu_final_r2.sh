#!/bin/bash
# final bench lines of round 2 (one visit): 1 M and 16 M spheres with the CPU arm, and the ncu capture of the 1 M-sphere kernel
set -u
P=r02b
mkdir -p gpurun_out
timeout 900 python bench.py --workload c4 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/${P}_bench_c4.json
timeout 1200 python bench.py --workload c5 --steps 4 --warmup 3 2>&1 | tail -1 > gpurun_out/${P}_bench_c5.json
for f in c4 c5; do python -c "
import json
d=json.loads(open('gpurun_out/${P}_bench_$f.json').read().strip().splitlines()[-1]); print('$f: %.0f Mrays/s e2e %.0f ms/step %.3f cpu %s roofline %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d.get('cpu_baseline',{}).get('value'), d['roofline']['frac']))"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_lean -s 4 -c 1 -f -o gpurun_out/${P}_prof_c4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/${P}_ncu_c4.log 2>&1
ncu -i gpurun_out/${P}_prof_c4.ncu-rep --page raw --csv > gpurun_out/${P}_raw_c4.csv 2>/dev/null
ncu -i gpurun_out/${P}_prof_c4.ncu-rep --page source --csv > gpurun_out/${P}_src_c4.csv 2>/dev/null
ls -la gpurun_out/${P}_*
