#!/bin/bash
# single-GPU evidence for the session (files land in gpurun_out/ with the prefix given as $1, default s3): full GPU test suite,
# smoke, the bench line (K=64), the reference arm, one line per kernel variant / BASELINE config, the ncu launch list and one
# ncu --set full capture of the headline kernel.
set -u
P=${1:-s3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${P}_gpu.txt
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/${P}_pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${P}_smoke.log
echo "=== bench K=64"
timeout 900 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-400
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 16 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${P}_bench_reference_n1.json | cut -c1-200
B="timeout 900 python bench.py --warmup 3 --no-cpu-baseline"
$B --steps 16 --opt accel=0 2>&1 | tail -1 > gpurun_out/${P}_bench_grid.json
$B --steps 16 --opt huge_factor=0 2>&1 | tail -1 > gpurun_out/${P}_bench_nohuge.json
$B --steps 16 --opt async_done=0 2>&1 | tail -1 > gpurun_out/${P}_bench_persistent.json
$B --steps 16 --opt tile_order=0 2>&1 | tail -1 > gpurun_out/${P}_bench_rowmajor.json
$B --steps 16 --opt warp_tiles=0 2>&1 | tail -1 > gpurun_out/${P}_bench_lane_tickets.json
$B --steps 16 --opt async_done=24 --opt async_node=8 2>&1 | tail -1 > gpurun_out/${P}_bench_async_voted.json
$B --steps 16 --opt wide_threads=768 2>&1 | tail -1 > gpurun_out/${P}_bench_768.json
$B --steps 16 --opt wide_nodes=0 --opt sah_max_prims=0 --leaf-size 2 2>&1 | tail -1 > gpurun_out/${P}_bench_karras_pairs.json
$B --steps 16 --workload c1 2>&1 | tail -1 > gpurun_out/${P}_bench_c1.json
$B --steps 8 --workload c3 2>&1 | tail -1 > gpurun_out/${P}_bench_c3_n1.json
$B --steps 4 --workload c4 2>&1 | tail -1 > gpurun_out/${P}_bench_c4.json
$B --steps 4 --workload c5 2>&1 | tail -1 > gpurun_out/${P}_bench_c5.json
for f in grid nohuge persistent rowmajor lane_tickets async_voted 768 karras_pairs c1 c3_n1 c4 c5; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/${P}_bench_$f.json').read().strip().splitlines()[-1]); print('$f: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('$f FAILED', e)
"; done
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_launches.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_async -s 4 -c 1 -f -o gpurun_out/${P}_prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_full.log 2>&1
ls -la gpurun_out/${P}_prof_final.ncu-rep
echo "=== launch timeline"
python tools/tail_probe.py 2>&1 | cut -c1-150 | tee gpurun_out/${P}_tail_probe.txt
