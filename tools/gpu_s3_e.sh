#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python tools/sweep_options.py \
  "async_done=0" "async_done=26,async_node=0" "async_done=24,async_node=0" "async_done=28,async_node=0" "async_done=0" 2>&1 | tee gpurun_out/sweep_e.log
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-700
