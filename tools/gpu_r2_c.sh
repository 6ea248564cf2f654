#!/bin/bash
# round 2, visit C: top-of-stack register in k_render_lean
set -u
P=${1:-r2c}
mkdir -p gpurun_out
echo "=== pytest subset"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "async_kernel or cost_ordered or default_options or rtiow_full_size" 2>&1 | tail -8 | tee gpurun_out/${P}_pytest_subset.log
echo "=== sweep"
timeout 600 python tools/sweep_options.py "lean=0" "lean=1" "lean=1,async_done=28" "lean=1,async_done=22" "lean=1,tile_order=3" "lean=1,tile_order=2" "lean=1,tile_order=1" 2>&1 | tee gpurun_out/${P}_sweep.log
echo "=== bench"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-300
echo "=== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_lean -s 4 -c 1 -f -o gpurun_out/${P}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_full.log 2>&1
ncu -i gpurun_out/${P}_prof.ncu-rep --page raw --csv > gpurun_out/${P}_raw.csv 2>/dev/null
ncu -i gpurun_out/${P}_prof.ncu-rep --page source --csv > gpurun_out/${P}_src.csv 2>/dev/null
ls -la gpurun_out/${P}_*
