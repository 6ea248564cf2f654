#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest (octant/wide/slot/c1)"
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "slot or wide or octant or c1_exact or tolerance" 2>&1 | tail -6
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s ms/step %.3f' % (d['value'], d['ms_per_step']))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
for LEAF in 1 2 3; do for V in 0 2 4 6 8 12 16; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --leaf-size $LEAF --opt leaf_vote=$V 2>&1 | show "leaf$LEAF vote=$V"
done; done
