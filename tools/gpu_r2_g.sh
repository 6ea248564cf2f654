#!/bin/bash
# round 2, visit G (1 GPU): full GPU suite
set -u
P=${1:-r2g}
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^\s*$" | tail -70 | tee gpurun_out/${P}_pytest_gpu.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-300
ls -la gpurun_out/${P}_*
