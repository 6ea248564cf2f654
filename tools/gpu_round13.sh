#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s e2e %.0f ms/step %.3f build %.2f ms nodes %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['bvh_build_ms'], d['config']['bvh_nodes']), d['roofline'].get('model','')[60:130])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
timeout 600 python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | show "c4 persistent"
timeout 900 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | show "c5 persistent"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes_read.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_render_persistent -s 3 -c 1 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "k_render|gpu__time|issue_active|lts__|dram__|smsp__|l1tex" 
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes_read.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_render_persistent -s 3 -c 1 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "k_render|gpu__time|issue_active|lts__|dram__|smsp__|l1tex" 
