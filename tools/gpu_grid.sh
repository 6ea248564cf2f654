#!/bin/bash
# uniform grid + oversize list: tests, A/B against the BVH, threads / vote sweeps, ncu
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s e2e %.0f ms/step %.3f [%s]' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config'].get('accel')), d['roofline'].get('model','')[:110])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline"
$B 2>&1 | show "default"
$B --opt accel=1 2>&1 | show "bvh"
for T in 768 512; do $B --opt wide_threads=$T 2>&1 | show "grid threads=$T"; done
for V in 0 4 8 16 24; do $B --opt leaf_vote=$V 2>&1 | show "grid vote=$V"; done
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_grid python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_grid.log 2>&1
ls -la gpurun_out/prof_grid.ncu-rep
