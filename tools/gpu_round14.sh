#!/bin/bash
set -u
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s ms/step %.3f' % (d['value'], d['ms_per_step']))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
for V in 0 4 8 12 16 24; do
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --opt leaf_vote=$V 2>&1 | show "c4 vote=$V"
done
for V in 0 12; do for T in 128 256; do
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --opt leaf_vote=$V --threads $T 2>&1 | show "c4 vote=$V threads=$T"
done; done
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --opt wide_nodes=0 --opt leaf_vote=12 --leaf-size 3 2>&1 | show "rtiow pairs-oct leaf3 vote=12"
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --opt wide_nodes=0 --opt leaf_vote=0 --leaf-size 3 2>&1 | show "rtiow pairs-oct leaf3 vote=0"
