#!/bin/bash
# last visit of the session: full GPU suite, smoke, the bench line, one ncu --set full capture of the default kernel, one A/B line
set -u
P=${1:-s3c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${P}_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${P}_smoke.log
timeout 200 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-330
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_render_async -s 4 -c 1 -f -o gpurun_out/${P}_prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_full.log 2>&1
timeout 100 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --opt warp_tiles=0 2>&1 | tail -1 > gpurun_out/${P}_bench_lane_tickets.json
timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload c3 2>&1 | tail -1 > gpurun_out/${P}_bench_c3_n1.json
ls -la gpurun_out/${P}_*
