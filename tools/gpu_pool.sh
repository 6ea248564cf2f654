#!/bin/bash
# pool-kernel bring-up: parity tests for the pool kernel, then a parameter sweep with short benches
set -u
mkdir -p gpurun_out
echo "=== pytest pool"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s -k "pool or empty or traversal or 1024spp" 2>&1 | tail -30 | tee gpurun_out/pytest_pool.log
echo "=== sweep"
for cfg in "96 768 8 8" "96 768 4 8" "96 768 12 8" "96 768 8 4" "96 768 8 16" "64 768 8 8" "64 512 8 8" "128 512 8 8" "160 384 8 8" "96 768 16 12" "96 768 1 1"; do
  set -- $cfg
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --kernel pool --pool-slots $1 --pool-threads $2 --pool-service $3 --pool-leaf-batch $4 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('pool slots=$1 threads=$2 service=$3 leaf=$4  value %.0f Mrays/s  ms/step %.2f' % (d['value'], d['ms_per_step']))
except Exception as e: print('pool $cfg FAILED', l[-300:])
" | tee -a gpurun_out/pool_sweep.log
done
for ls in 1 2 4 8; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --leaf-size $ls 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip()); print('persistent leaf_size=$ls value %.0f Mrays/s' % d['value'])" | tee -a gpurun_out/pool_sweep.log
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --kernel pool --leaf-size $ls 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip()); print('pool leaf_size=$ls value %.0f Mrays/s' % d['value'])" | tee -a gpurun_out/pool_sweep.log
done
