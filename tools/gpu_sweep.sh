#!/bin/bash
# option sweep of the headline workload within one visit: gpu_sweep.sh <prefix> "<opts 1>" "<opts 2>" ...   (each = space-separated key=value list)
set -u
P=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for o in "$@"; do
args=""; for kv in $o; do args="$args --opt $kv"; done
timeout 600 python bench.py --warmup 3 --no-cpu-baseline --strong-subframes 0 --steps 32 $args 2>&1 | tail -1 > gpurun_out/${P}_bench.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench.json').read()); print('[$o]: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done; done
