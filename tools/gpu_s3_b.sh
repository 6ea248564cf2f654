#!/bin/bash
# session 3, visit B: parity of the normalised/saturating node step (PTX control block), rnd_pm1, sphere_root shortcuts; sweep
set -u
mkdir -p gpurun_out
echo "=== pytest subset"
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x \
  -k "async_kernel or octant_and_plain or c1_exact_build or default_options or wide_nodes_match or exact_build_other or traversal_equals_brute or rng_device or scatter_device or empty_single or baseline_tolerance or slot_kernel or grid_and_bvh" 2>&1 | tail -15 | tee gpurun_out/pytest_subset_b.log
echo "=== sweep"
timeout 600 python tools/sweep_options.py \
  "async_done=0,wide_threads=768" \
  "async_done=0,wide_threads=1024" \
  "async_done=0,wide_threads=512" \
  "async_done=24,async_node=8,async_leaf=8,wide_threads=1024" \
  "async_done=24,async_node=8,async_leaf=8,wide_threads=768" \
  "async_done=28,async_node=8,async_leaf=8,wide_threads=1024" \
  "async_done=20,async_node=8,async_leaf=8,wide_threads=1024" \
  "async_done=24,async_node=4,async_leaf=4,wide_threads=1024" \
  "async_done=24,async_node=12,async_leaf=6,wide_threads=1024" \
  "async_done=32,async_node=8,async_leaf=8,wide_threads=1024" \
  "async_done=0,wide_threads=1024,leaf_vote=12" \
  "async_done=0,leaf_vote=0,wide_threads=1024" 2>&1 | tee gpurun_out/sweep_b.log
echo "=== ncu full"
SWEEP_FRAMES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 2 -c 1 -f -o gpurun_out/prof_s3b_persistent python tools/sweep_options.py "async_done=0,wide_threads=1024" > gpurun_out/ncu_persistent_b.log 2>&1
SWEEP_FRAMES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_async -s 2 -c 1 -f -o gpurun_out/prof_s3b_async python tools/sweep_options.py "async_done=24,async_node=8,async_leaf=8,wide_threads=1024" > gpurun_out/ncu_async_b.log 2>&1
ls -la gpurun_out/ | tail -8
