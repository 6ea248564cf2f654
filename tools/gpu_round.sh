#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + one full capture.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 600 python bench.py --steps ${BENCH_STEPS:-64} --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "=== bench --fast"
timeout 600 python bench.py --steps 16 --warmup 3 --fast --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_fast.log
if [ "${WITH_WAVEFRONT:-1}" = "1" ]; then
  echo "=== bench wavefront"
  timeout 600 python bench.py --steps 8 --warmup 3 --kernel wavefront --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_wavefront.log
fi
if [ "${WITH_NCU:-1}" = "1" ]; then
  echo "=== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
  echo "=== ncu full capture of the path kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_path \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
  ls -la gpurun_out/
fi
