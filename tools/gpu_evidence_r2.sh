#!/bin/bash
# round 2 evidence (1 GPU): full GPU suite, smoke, the bench line (K=64) + reference arm, the other configs, ncu launch list,
# ncu --set full of the headline kernel and of the L2/HBM kernel on configs[3] / configs[4].  Files: gpurun_out/<prefix>_*
set -u
P=${1:-r2ev}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${P}_gpu.txt
echo "=== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/${P}_pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${P}_smoke.log
echo "=== bench K=64"
timeout 900 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-300
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 16 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${P}_bench_reference_n1.json | cut -c1-200
B="timeout 900 python bench.py --warmup 3 --no-cpu-baseline"
$B --steps 16 --opt lean=0 2>&1 | tail -1 > gpurun_out/${P}_bench_async_round1_kernel.json
$B --steps 16 --workload c1 2>&1 | tail -1 > gpurun_out/${P}_bench_c1.json
$B --steps 8 --workload c3 2>&1 | tail -1 > gpurun_out/${P}_bench_c3_n1.json
$B --steps 4 --workload c4 2>&1 | tail -1 > gpurun_out/${P}_bench_c4.json
$B --steps 4 --workload c5 2>&1 | tail -1 > gpurun_out/${P}_bench_c5.json
$B --steps 4 --workload c4 --opt lean=0 2>&1 | tail -1 > gpurun_out/${P}_bench_c4_persistent.json
$B --steps 8 --kernel wavefront 2>&1 | tail -1 > gpurun_out/${P}_bench_wavefront.json
for f in async_round1_kernel c1 c3_n1 c4 c5 c4_persistent wavefront; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/${P}_bench_$f.json').read().strip().splitlines()[-1]); print('$f: %.0f Mrays/s e2e %.0f ms/step %.3f build %.3f ms' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['bvh_build']['ms']))
except Exception as e: print('$f FAILED', e)
"; done
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-subframes 0 > gpurun_out/${P}_ncu_launches.log 2>&1
echo "=== ncu full (headline)"
# (one launch per frame for the capture: with split frames every second launch of the kernel is the cheap sky part)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_lean -s 4 -c 1 -f -o gpurun_out/${P}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-subframes 0 --opt split_tail=0 > gpurun_out/${P}_ncu_full.log 2>&1
ncu -i gpurun_out/${P}_prof.ncu-rep --page raw --csv > gpurun_out/${P}_raw.csv 2>/dev/null
ncu -i gpurun_out/${P}_prof.ncu-rep --page source --csv > gpurun_out/${P}_src.csv 2>/dev/null
for c in c4 c5; do
echo "=== ncu full ($c)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_render_lean -s 4 -c 1 -f -o gpurun_out/${P}_prof_$c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $c > gpurun_out/${P}_ncu_$c.log 2>&1
ncu -i gpurun_out/${P}_prof_$c.ncu-rep --page raw --csv > gpurun_out/${P}_raw_$c.csv 2>/dev/null
ncu -i gpurun_out/${P}_prof_$c.ncu-rep --page source --csv > gpurun_out/${P}_src_$c.csv 2>/dev/null
done
echo "=== throughput over time (instrumented launches)"
timeout 300 python tools/tail_probe.py profile=1 only=1080 > gpurun_out/${P}_throughput_over_time.txt 2>&1
timeout 600 python tools/tail_probe.py scene=1m units=1 steal=0 | grep "launch [12]:" > gpurun_out/${P}_drain_1m.txt 2>&1
timeout 600 python tools/tail_probe.py scene=1m units=1 | grep "launch [12]:" >> gpurun_out/${P}_drain_1m.txt 2>&1
timeout 600 python tools/tail_probe.py scene=1m | grep "launch [12]:" >> gpurun_out/${P}_drain_1m.txt 2>&1
VN_DEBUG_SPLIT=1 timeout 300 python tools/split_probe.py "split_tail=0" "split_tail=0.5" > gpurun_out/${P}_split_probe.txt 2>&1
timeout 300 python tools/guess_probe.py > gpurun_out/${P}_guess_probe.txt 2>&1
echo "=== drain"
timeout 200 python tools/tail_probe.py 2>&1 | grep "launch" | tee gpurun_out/${P}_tail_probe.txt
ls -la gpurun_out/${P}_* | awk '{print $5, $9}'
