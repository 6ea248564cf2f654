#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: value %.0f Mrays/s  e2e %.0f  ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('$1 FAILED', l[-300:])
"; }
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_oct.log | show "octant(default)"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --opt octant_nodes=0 2>&1 | tee gpurun_out/bench_nooct.log | show "plain 256x4"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --opt octant_nodes=0 --threads 128 2>&1 | show "plain 128"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --leaf-size 1 2>&1 | show "octant leaf1"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --leaf-size 4 2>&1 | show "octant leaf4"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --fast 2>&1 | show "octant VN_FAST"
