#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== bvh build"
timeout 900 python tools/bvh_build_bench.py 1000000 16000000 2>&1 | tee gpurun_out/bvh_build.log
echo "=== bench"
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
