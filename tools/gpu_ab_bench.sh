#!/bin/bash
# A/B of compile-time variants with bench.py (L2 flush between steps): gpu_ab_bench.sh "<exp numbers>" <workload> <steps> [bench args]
set -u
EXPS=$1; W=${2:-c2}; K=${3:-32}; shift 3
mkdir -p gpurun_out
for rep in 1 2; do
for e in 0 $EXPS; do
if [ $e = 0 ]; then unset VN_EXPERIMENT; else export VN_EXPERIMENT=$e; fi
timeout 900 python bench.py --warmup 3 --no-cpu-baseline --strong-subframes 0 --workload $W --steps $K "$@" 2>&1 | tail -1 > gpurun_out/ab_bench.json
python -c "
import json; d=json.loads(open('gpurun_out/ab_bench.json').read()); print('$W exp $e: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done; done
