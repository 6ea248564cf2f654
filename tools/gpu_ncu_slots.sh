#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_slots -s 4 -c 1 -f -o gpurun_out/prof_slots python bench.py --steps 2 --warmup 3 --no-cpu-baseline --kernel slots --leaf-size ${LEAF:-3} ${EXTRA:-} > gpurun_out/ncu_full_slots.log 2>&1
ls -la gpurun_out/prof_slots.ncu-rep
