#!/bin/bash
# session 3, visit A: parity of the async kernel + FFMA2 node step, threshold sweep, ncu full captures of both kernels
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== pytest subset"
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x \
  -k "async_kernel or octant_and_plain or c1_exact_build or default_options or wide_nodes_match or exact_build_other" 2>&1 | tail -15 | tee gpurun_out/pytest_subset.log
echo "=== sweep"
timeout 600 python tools/sweep_options.py \
  "async_done=0" \
  "async_done=32,async_node=8,async_leaf=8" \
  "async_done=28,async_node=8,async_leaf=8" \
  "async_done=24,async_node=8,async_leaf=8" \
  "async_done=20,async_node=8,async_leaf=8" \
  "async_done=16,async_node=8,async_leaf=8" \
  "async_done=24,async_node=1,async_leaf=1" \
  "async_done=24,async_node=4,async_leaf=8" \
  "async_done=24,async_node=12,async_leaf=8" \
  "async_done=24,async_node=16,async_leaf=12" \
  "async_done=24,async_node=8,async_leaf=4" \
  "async_done=24,async_node=8,async_leaf=12" \
  "async_done=28,async_node=12,async_leaf=12" \
  "async_done=24,async_node=8,async_leaf=8,wide_threads=512" \
  "async_done=0,wide_threads=768,leaf_vote=12" \
  "async_done=0,leaf_vote=0,wide_threads=1024" \
  "async_done=0,wide_threads=768" 2>&1 | tee gpurun_out/sweep_a.log
echo "=== ncu full: persistent (default) and async"
SWEEP_FRAMES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 2 -c 1 -f -o gpurun_out/prof_s3_persistent python tools/sweep_options.py "async_done=0" > gpurun_out/ncu_persistent.log 2>&1
SWEEP_FRAMES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_async -s 2 -c 1 -f -o gpurun_out/prof_s3_async python tools/sweep_options.py "async_done=24,async_node=8,async_leaf=8" > gpurun_out/ncu_async.log 2>&1
ls -la gpurun_out/
