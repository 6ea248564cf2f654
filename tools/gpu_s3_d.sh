#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "async_kernel" 2>&1 | tail -5
timeout 600 python tools/sweep_options.py \
  "async_done=0,wide_threads=1024" \
  "async_done=31,async_node=0,wide_threads=1024" \
  "async_done=30,async_node=0,wide_threads=1024" \
  "async_done=28,async_node=0,wide_threads=1024" \
  "async_done=26,async_node=0,wide_threads=1024" \
  "async_done=24,async_node=0,wide_threads=1024" \
  "async_done=20,async_node=0,wide_threads=1024" \
  "async_done=28,async_node=0,wide_threads=768" \
  "async_done=0,wide_threads=1024" 2>&1 | tee gpurun_out/sweep_d.log
SWEEP_FRAMES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_async -s 2 -c 1 -f -o gpurun_out/prof_s3d_phase python tools/sweep_options.py "async_done=28,async_node=0,wide_threads=1024" > gpurun_out/ncu_phase.log 2>&1
