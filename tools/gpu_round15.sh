#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest (octant/wide/c1/tolerance/progressive)"
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "wide or octant or c1_exact or tolerance or progressive or launch_shapes" 2>&1 | tail -6
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s ms/step %.3f' % (d['value'], d['ms_per_step']))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
for D in 0 12 16 20 24 28; do for V in 8 12 16; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --opt leaf_vote=$V --opt done_vote=$D 2>&1 | show "leaf1 leaf_vote=$V done_vote=$D"
done; done
for LEAF in 2 3; do for D in 0 20; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --leaf-size $LEAF --opt done_vote=$D 2>&1 | show "leaf$LEAF leaf_vote=12 done_vote=$D"
done; done
