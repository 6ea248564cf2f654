#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/sweep_options.py "tile_order=0" "tile_order=1" "tile_order=2" "tile_order=3" "tile_order=1" "tile_order=3" 2>&1 | tee gpurun_out/sweep_g.log
python tools/tail_probe.py 2>&1 | cut -c1-140
