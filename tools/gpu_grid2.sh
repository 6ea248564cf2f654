#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s e2e %.0f ms/step %.3f [%s]' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config'].get('accel')))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline"
$B 2>&1 | show "default (bvh)"
for V in 0 1 2 3 4 6; do $B --opt accel=0 --opt grid_vote=$V 2>&1 | show "grid vote=$V"; done
$B --opt accel=0 --opt grid_vote=0 --opt wide_threads=512 2>&1 | show "grid vote=0 threads=512"
$B --opt accel=0 --opt grid_vote=4 --opt wide_threads=512 2>&1 | show "grid vote=4 threads=512"
