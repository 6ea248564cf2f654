#!/bin/bash
# round 2, visit E (1 GPU): full GPU suite, smoke, bench line (new bench.py), c4 / c5 with 256-bit pair loads
set -u
P=${1:-r2e}
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -12 | tee gpurun_out/${P}_pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${P}_smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-400
echo "=== sweep c4"
timeout 900 python tools/sweep_large.py c4 "lean=1" "lean=1,global_done=12" "lean=1,global_done=20" "lean=1,global_done=16,async_leaf=4" 2>&1 | tee gpurun_out/${P}_sweep_c4.log
echo "=== sweep c5"
timeout 1200 python tools/sweep_large.py c5 "lean=1" "lean=1,global_done=12" 2>&1 | tee gpurun_out/${P}_sweep_c5.log
echo "=== bench c4"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload c4 2>&1 | tail -1 | tee gpurun_out/${P}_bench_c4.json | cut -c1-300
ls -la gpurun_out/${P}_*
