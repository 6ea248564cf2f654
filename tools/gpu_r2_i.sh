#!/bin/bash
set -u
P=${1:-r2i}
mkdir -p gpurun_out
echo "=== sweep tile_order"
timeout 600 python tools/sweep_options.py "tile_order=1" "tile_order=4" "tile_order=1" "tile_order=4" "tile_order=4,async_done=24" "tile_order=4,async_done=28" 2>&1 | tee gpurun_out/${P}_sweep.log
echo "=== drain mode 4"
timeout 300 python tools/tail_probe.py tile_order=4 2>&1 | head -40 | tee gpurun_out/${P}_tail_probe.txt
