#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu (full)"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 2>&1 | tail -22 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "=== bench default"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f ms/step %.3f frac %.3f cpu %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value']))"
