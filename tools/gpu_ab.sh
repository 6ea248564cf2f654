#!/bin/bash
# A/B of compile-time variants within one visit: the shipped library and the experiment libraries (python -m venusaur_b200.build --exp N)
# usage: gpu_ab.sh <prefix> "<exp numbers>" [split_probe option sets...]
set -u
P=$1; EXPS=$2; shift 2
for rep in 1 2; do
for e in 0 $EXPS; do
if [ $e = 0 ]; then unset VN_EXPERIMENT; else export VN_EXPERIMENT=$e; fi
echo "--- library: experiment $e"
python tools/split_probe.py "$@" 2>&1 | grep "ms_render"
done; done
