#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: value %.0f Mrays/s  e2e %.0f  ms/step %.3f build_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['bvh_build_ms']), d['roofline']['model'][:100])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | show "sah leaf2 (default)"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --leaf-size 4 2>&1 | show "sah leaf4"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --leaf-size 3 2>&1 | show "sah leaf3"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --leaf-size 1 2>&1 | show "sah leaf1"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --opt sah_max_prims=0 2>&1 | show "karras leaf2"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --opt aabb_pad=0.002 2>&1 | show "sah leaf2 pad0.2%"
