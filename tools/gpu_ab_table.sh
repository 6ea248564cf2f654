#!/bin/bash
# A/B table of the persistent kernel: split heuristic x leaf size x node layout (plain lo/hi, 8 octant copies, 4-wide octant-sorted)
set -u
mkdir -p gpurun_out
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s  ms/step %.3f nodes %d' % (d['value'], d['ms_per_step'], d['config']['bvh_nodes']), d['roofline'].get('model','')[60:130])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
for SAH in 4096 0; do for LEAF in 1 2 3 4; do
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --leaf-size $LEAF --opt sah_max_prims=$SAH"
$B --opt wide_nodes=0 --opt octant_nodes=0 2>&1 | show "sah=$SAH leaf$LEAF plain256"
$B --opt wide_nodes=0 --opt octant_nodes=0 --threads 512 2>&1 | show "sah=$SAH leaf$LEAF plain512"
$B --opt wide_nodes=0 --opt octant_nodes=1 2>&1 | show "sah=$SAH leaf$LEAF octant"
$B --opt wide_nodes=1 2>&1 | show "sah=$SAH leaf$LEAF wide"
done; done
