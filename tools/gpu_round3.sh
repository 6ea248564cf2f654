#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "=== bench c2 K=64"
timeout 600 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
echo "=== ncu full (octant persistent kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_path python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
echo "=== c4 (1M spheres)"
for k in persistent wavefront pool; do
  timeout 600 python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline --kernel $k 2>&1 | tail -1 | tee gpurun_out/bench_c4_$k.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('c4 $k: value %.0f Mrays/s ms/step %.2f build_ms %.2f seg/path %.2f' % (d['value'], d['ms_per_step'], d['config']['bvh_build_ms'], d['segments_per_path']), d['roofline']['model'][:120])
except Exception as e: print('c4 $k FAILED', l[-400:])
"
done
ls -la gpurun_out | head -30
