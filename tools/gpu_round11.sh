#!/bin/bash
# new default (SAH splits, 4-wide octant-sorted nodes, auto leaf size, PTX-predicated pushes): full test suite, bench, profiles
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() { python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('$1: %.0f Mrays/s e2e %.0f ms/step %.3f nodes %d leaf %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['bvh_nodes'], d['config']['leaf_size']), d['roofline'].get('model','')[60:130])
except Exception as e: print('$1 FAILED', l[-300:])
"; }
B="timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline"
$B 2>&1 | show "default"
$B --leaf-size 2 2>&1 | show "leaf2"
$B --leaf-size 3 2>&1 | show "leaf3"
echo "=== bench K=64 (full line)"
timeout 900 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_wide python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_wide.log 2>&1
ls -la gpurun_out/prof_wide.ncu-rep
