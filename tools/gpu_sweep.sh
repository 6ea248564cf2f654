#!/bin/bash
# option sweep of a bench workload within one visit: gpu_sweep.sh <workload> <steps> "<opts 1>" "<opts 2>" ...   (each = space-separated key=value list)
set -u
W=$1; K=$2; shift 2
mkdir -p gpurun_out
for o in "$@"; do
args=""; for kv in $o; do args="$args --opt $kv"; done
timeout 900 python bench.py --warmup 3 --no-cpu-baseline --strong-subframes 0 --workload $W --steps $K $args 2>&1 | tail -1 > gpurun_out/sweep_bench.json
python -c "
import json; d=json.loads(open('gpurun_out/sweep_bench.json').read()); print('$W [$o]: %.0f Mrays/s e2e %.0f ms/step %.3f build %.2f ms' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['bvh_build']['ms']))"
done
