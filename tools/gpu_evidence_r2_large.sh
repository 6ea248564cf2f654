#!/bin/bash
set -u
P=r2fin2
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${P}_pytest_gpu.log
B="timeout 900 python bench.py --warmup 3 --no-cpu-baseline"
$B --steps 4 --workload c4 2>&1 | tail -1 > gpurun_out/${P}_bench_c4.json
$B --steps 4 --workload c5 2>&1 | tail -1 > gpurun_out/${P}_bench_c5.json
for c in c4 c5; do
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_render_lean -s 4 -c 1 -f -o gpurun_out/${P}_prof_$c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $c > gpurun_out/${P}_ncu_$c.log 2>&1
ncu -i gpurun_out/${P}_prof_$c.ncu-rep --page raw --csv > gpurun_out/${P}_raw_$c.csv 2>/dev/null
ncu -i gpurun_out/${P}_prof_$c.ncu-rep --page source --csv > gpurun_out/${P}_src_$c.csv 2>/dev/null
done
timeout 600 python tools/tail_probe.py scene=1m | grep "launch [12]:" > gpurun_out/${P}_drain_1m.txt 2>&1
for f in c4 c5; do python -c "
import json
d=json.loads(open('gpurun_out/${P}_bench_$f.json').read().strip().splitlines()[-1]); print('$f: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; done
cat gpurun_out/${P}_drain_1m.txt
