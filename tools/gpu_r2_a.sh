#!/bin/bash
# round 2, visit A (code as of round 1): ncu evidence for the L2/HBM trace kernel on configs[3] / configs[4] and for the LBVH builder
# kernels at 1 M and 16 M spheres, plus the baseline numbers this round starts from.  Files land in gpurun_out/r2a_*.
set -u
P=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${P}_gpu.txt
echo "=== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${P}_pytest_gpu.log
echo "=== bench c2"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-300
echo "=== builder timing"
timeout 600 python tools/bvh_build_bench.py 1000000 16000000 2>&1 | tee gpurun_out/${P}_bvh_build.txt
echo "=== builder kernels under ncu (time + dram bytes)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${P}_builder_launches.csv python tools/bvh_build_bench.py 1000000 16000000 > gpurun_out/${P}_builder_ncu.log 2>&1
echo "=== ncu full c4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/${P}_prof_c4 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/${P}_ncu_c4.log 2>&1
echo "=== ncu full c5"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 4 -c 1 -f -o gpurun_out/${P}_prof_c5 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c5 > gpurun_out/${P}_ncu_c5.log 2>&1
echo "=== bench c4 / c5"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload c4 2>&1 | tail -1 | tee gpurun_out/${P}_bench_c4.json | cut -c1-200
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload c5 2>&1 | tail -1 | tee gpurun_out/${P}_bench_c5.json | cut -c1-200
echo "=== drain"
timeout 200 python tools/tail_probe.py 2>&1 | tee gpurun_out/${P}_tail_probe.txt
ls -la gpurun_out/${P}_*
