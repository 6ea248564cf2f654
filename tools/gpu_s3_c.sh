#!/bin/bash
# session 3, visit C: full GPU suite after the material-record change + sweep incl. the grid
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu (full)"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=6 2>&1 | tail -20 | tee gpurun_out/pytest_gpu_c.log
echo "=== sweep"
timeout 600 python tools/sweep_options.py \
  "async_done=0,wide_threads=1024" \
  "async_done=0,wide_threads=768" \
  "async_done=24,async_node=8,async_leaf=8,wide_threads=1024" \
  "async_done=0,wide_threads=1024,accel=0" \
  "async_done=0,wide_threads=768,accel=0" \
  "async_done=0,wide_threads=1024,accel=1,huge_factor=0" \
  "async_done=0,wide_threads=1024,accel=1,huge_factor=50" 2>&1 | tee gpurun_out/sweep_c.log
