#!/bin/bash
# A/B of compile-time variants within one visit: the shipped library and the experiment libraries (python -m venusaur_b200.build --exp N)
# usage: gpu_ab.sh <prefix> "<exp numbers>" [bench args]
set -u
P=$1; EXPS=$2; shift 2
mkdir -p gpurun_out
B="timeout 600 python bench.py --warmup 3 --no-cpu-baseline --steps 32 --strong-subframes 0 $*"
for rep in 1 2; do
for e in 0 $EXPS; do
if [ $e = 0 ]; then unset VN_EXPERIMENT; else export VN_EXPERIMENT=$e; fi
$B 2>&1 | tail -1 > gpurun_out/${P}_bench_exp$e.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_exp$e.json').read()); print('exp $e: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done; done
