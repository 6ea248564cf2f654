#!/bin/bash
set -u
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
for T in 1024 768 512; do
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --opt wide_threads=$T 2>&1 | show "wide threads=$T"
done
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --opt wide_threads=768 --opt leaf_vote=8 2>&1 | show "wide threads=768 vote 8"
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --opt wide_threads=768 --opt leaf_vote=16 2>&1 | show "wide threads=768 vote 16"
