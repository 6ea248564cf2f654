#!/bin/bash
# round 2, visit F (1 GPU): full GPU suite (all failures), smoke, bench line, c4 / c5 with the hit-point gate
set -u
P=${1:-r2f}
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^\s*$" | tail -60 | tee gpurun_out/${P}_pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${P}_smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-300
echo "=== sweep c4 / c5"
timeout 900 python tools/sweep_large.py c4 "lean=1" "lean=1,hit_gate=0" 2>&1 | tee gpurun_out/${P}_sweep_c4.log
timeout 1200 python tools/sweep_large.py c5 "lean=1" 2>&1 | tee gpurun_out/${P}_sweep_c5.log
ls -la gpurun_out/${P}_*
