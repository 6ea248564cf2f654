#!/bin/bash
# round 2, visit H (1 GPU): what is in the drain (lane-retirement histogram), CTAs per SM for the L2/HBM kernel
set -u
P=${1:-r2h}
mkdir -p gpurun_out
echo "=== drain"
timeout 300 python tools/tail_probe.py 2>&1 | tee gpurun_out/${P}_tail_probe.txt
echo "=== sweep c4"
timeout 900 python tools/sweep_large.py c4 "global_ctas=4" "global_ctas=5" "global_ctas=6" "global_ctas=5,global_done=20" "global_ctas=5,global_done=12" 2>&1 | tee gpurun_out/${P}_sweep_c4.log
echo "=== sweep c5"
timeout 1200 python tools/sweep_large.py c5 "global_ctas=4" "global_ctas=5" "global_ctas=6" 2>&1 | tee gpurun_out/${P}_sweep_c5.log
