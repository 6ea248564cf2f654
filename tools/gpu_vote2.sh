#!/bin/bash
set -u
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s ms/step %.3f' % (d['value'], d['ms_per_step']))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline"
for V in 0 2 4 6 12 0 12; do $B --opt leaf_vote=$V 2>&1 | show "leaf_vote=$V"; done
$B --opt leaf_vote=0 --opt wide_threads=768 2>&1 | show "leaf_vote=0 threads=768"
$B --opt leaf_vote=0 --leaf-size 2 2>&1 | show "leaf_vote=0 leaf2"
