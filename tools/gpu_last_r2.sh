#!/bin/bash
# last visit of round 2 (one GPU, ~3 min): the default bench line of the final code and the ncu launch list of the same command
set -u
P=r02d
mkdir -p gpurun_out
timeout 110 python bench.py 2>gpurun_out/${P}_bench_n1.err | tail -1 > gpurun_out/${P}_bench_n1.json
python -c "
import json
d=json.loads(open('gpurun_out/${P}_bench_n1.json').read().strip().splitlines()[-1]); print('n1: %.0f Mrays/s e2e %.0f ms/step %.3f one-call %.0f cpu %s roofline %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['subframes_in_one_call']['value'], d.get('cpu_baseline',{}).get('value'), d['roofline']['frac'], d['gpu_launches']))"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-subframes 0 > gpurun_out/${P}_ncu_launches.log 2>&1
tail -2 gpurun_out/${P}_ncu_launches.log | cut -c1-300
ls -la gpurun_out/${P}_*
