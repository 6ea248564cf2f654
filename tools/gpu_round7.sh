#!/bin/bash
# round-1 session 2: SAH + 4-wide nodes.  pytest, then A/B bench lines, then ncu of the new default kernel.
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: value %.0f Mrays/s  e2e %.0f  ms/step %.3f build_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['bvh_build_ms']), d['roofline'].get('model','')[:160])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline"
$B 2>&1 | tee gpurun_out/bench_r7_default.log | show "default (sah, wide, leaf auto)"
$B --leaf-size 2 2>&1 | show "sah wide leaf2"
$B --leaf-size 1 2>&1 | show "sah wide leaf1"
$B --leaf-size 4 2>&1 | show "sah wide leaf4"
$B --opt sah_max_prims=0 2>&1 | show "karras wide leaf2"
$B --opt wide_nodes=0 2>&1 | show "sah pairs-octant leaf3"
$B --opt wide_nodes=0 --opt sah_max_prims=0 2>&1 | show "karras pairs-octant leaf2 (old default)"
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_wide python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_wide.log 2>&1
ls -la gpurun_out/prof_wide.ncu-rep
