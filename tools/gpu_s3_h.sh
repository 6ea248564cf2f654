#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "async_kernel or cost_ordered or c1_exact_build or empty_single" 2>&1 | tail -8
timeout 200 python tools/sweep_options.py "warp_tiles=0" "warp_tiles=1" "warp_tiles=1,async_done=28" "warp_tiles=0" 2>&1 | tee gpurun_out/sweep_h.log
