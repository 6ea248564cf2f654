#!/bin/bash
# final single-GPU evidence for this session: full test suite, smoke, the bench line (K=64), reference arm, ncu launch list + full capture
# of the default kernel, and one bench line per other BASELINE config / kernel variant.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "=== bench K=64"
timeout 900 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 16 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-200
B="timeout 900 python bench.py --warmup 3 --no-cpu-baseline"
$B --steps 16 --opt accel=0 2>&1 | tail -1 > gpurun_out/bench_grid.log
$B --steps 16 --opt huge_factor=0 2>&1 | tail -1 > gpurun_out/bench_nohuge.log
$B --steps 16 --kernel slots --leaf-size 3 2>&1 | tail -1 > gpurun_out/bench_slots.log
$B --steps 16 --opt wide_nodes=0 --opt sah_max_prims=0 --leaf-size 2 2>&1 | tail -1 > gpurun_out/bench_karras_pairs.log
$B --steps 16 --opt leaf_vote=12 --opt wide_threads=1024 2>&1 | tail -1 > gpurun_out/bench_vote12.log
$B --steps 16 --workload c1 2>&1 | tail -1 > gpurun_out/bench_c1.log
$B --steps 8 --workload c3 2>&1 | tail -1 > gpurun_out/bench_c3_n1.log
$B --steps 4 --workload c4 2>&1 | tail -1 > gpurun_out/bench_c4.log
$B --steps 4 --workload c5 2>&1 | tail -1 > gpurun_out/bench_c5.log
for f in grid nohuge slots karras_pairs vote12 c1 c3_n1 c4 c5; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_$f.log').read().strip().splitlines()[-1]); print('$f: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('$f FAILED', e)
"; done
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
ls -la gpurun_out/prof_final.ncu-rep
