#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -6
timeout 600 python tools/sweep_options.py "tile_order=0" "tile_order=1" "tile_order=0,async_done=0" "tile_order=1,async_done=0" "tile_order=1,async_done=26" 2>&1 | tee gpurun_out/sweep_f.log
python tools/tail_probe.py 2>&1 | cut -c1-140
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload c3 2>&1 | tail -1 | cut -c1-200
