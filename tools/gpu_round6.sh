#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench K=64"
timeout 600 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_path python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out/prof_path.ncu-rep
