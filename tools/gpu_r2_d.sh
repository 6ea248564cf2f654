#!/bin/bash
# round 2, visit D: asynchronous path kernel over pair nodes in L2 / HBM (k_render_lean<kGlobal>) on configs[3] / configs[4]
set -u
P=${1:-r2d}
mkdir -p gpurun_out
echo "=== pytest (large scenes, config parity)"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -s -k "large_scene or config5 or lbvh_large or synthetic_scenes or empty_single" 2>&1 | tail -25 | tee gpurun_out/${P}_pytest_subset.log
echo "=== sweep c4"
timeout 900 python tools/sweep_large.py c4 "lean=0" "lean=1" "lean=1,async_done=16" "lean=1,async_done=8" "lean=1,async_done=16,async_leaf=4" "lean=1,async_done=16,async_leaf=16" "lean=1,async_done=30,async_leaf=8" 2>&1 | tee gpurun_out/${P}_sweep_c4.log
echo "=== sweep c5"
timeout 1200 python tools/sweep_large.py c5 "lean=0" "lean=1" "lean=1,async_done=16" "lean=1,async_done=8" "lean=1,async_done=16,async_leaf=4" 2>&1 | tee gpurun_out/${P}_sweep_c5.log
echo "=== ncu c4 (lean)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_lean -s 4 -c 1 -f -o gpurun_out/${P}_prof_c4 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/${P}_ncu_c4.log 2>&1
ncu -i gpurun_out/${P}_prof_c4.ncu-rep --page raw --csv > gpurun_out/${P}_raw_c4.csv 2>/dev/null
ncu -i gpurun_out/${P}_prof_c4.ncu-rep --page source --csv > gpurun_out/${P}_src_c4.csv 2>/dev/null
ls -la gpurun_out/${P}_*
