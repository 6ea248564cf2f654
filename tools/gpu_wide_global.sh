#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s ms/step %.3f build %.2f ms' % (d['value'], d['ms_per_step'], d['config']['bvh_build_ms']), d['roofline'].get('model','')[60:130])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
for WL in c4 c5; do
timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | show "$WL wide-global vote=12"
timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --opt leaf_vote=0 2>&1 | show "$WL wide-global vote=0"
timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --opt leaf_vote=4 2>&1 | show "$WL wide-global vote=4"
timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --opt wide_nodes=0 2>&1 | show "$WL pairs"
done
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --threads 128 2>&1 | show "c4 wide-global threads=128"
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --leaf-size 1 2>&1 | show "c4 wide-global leaf1"
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --leaf-size 4 2>&1 | show "c4 wide-global leaf4"
