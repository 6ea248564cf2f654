#!/bin/bash
set -u
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s ms/step %.3f' % (d['value'], d['ms_per_step']), d['roofline'].get('model','')[88:140])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline"
for F in 50 3 0; do $B --opt huge_factor=$F 2>&1 | show "huge_factor=$F"; done
for V in 6 8 10; do $B --opt leaf_vote=$V 2>&1 | show "leaf_vote=$V"; done
$B --opt leaf_vote=8 --opt wide_threads=768 2>&1 | show "leaf_vote=8 threads=768"
$B --opt huge_factor=3 --opt leaf_vote=8 2>&1 | show "huge_factor=3 leaf_vote=8"
