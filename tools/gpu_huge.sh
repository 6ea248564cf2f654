#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s e2e %.0f ms/step %.3f [%s]' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config'].get('accel')), d['roofline'].get('model','')[40:140])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline"
$B 2>&1 | show "default (bvh + huge list)"
$B --leaf-size 2 2>&1 | show "leaf2 (no huge list)"
for V in 8 16; do $B --opt leaf_vote=$V 2>&1 | show "default leaf_vote=$V"; done
$B --opt accel=0 2>&1 | show "grid"
