#!/bin/bash
# compute-sanitizer over tools/sanitize_probe.py (one GPU): gpu_sanitize.sh [tools...]   default: memcheck
set -u
mkdir -p gpurun_out
for tool in "${@:-memcheck}"; do
timeout -k 5 70 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 12 python tools/sanitize_probe.py > gpurun_out/r02d_sanitize_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|done in|RACECHECK SUMMARY" gpurun_out/r02d_sanitize_$tool.log | cut -c1-200
grep -E "=========     at |========= (Uninit|Invalid|Race|Barrier|Error)" gpurun_out/r02d_sanitize_$tool.log | sort | uniq -c | sort -rn | head -12 | cut -c1-260
done
