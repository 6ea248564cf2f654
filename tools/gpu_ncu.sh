#!/bin/bash
# ncu full capture of one kernel: usage gpu_ncu.sh <kernel-regex> <out-name> <bench args...>
set -u
mkdir -p gpurun_out
K=$1; OUT=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/$OUT python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${OUT}_bench.log 2>&1
tail -2 gpurun_out/${OUT}_bench.log | cut -c1-300
ls -la gpurun_out/$OUT.ncu-rep
