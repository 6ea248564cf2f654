#!/bin/bash
# round 2, visit n: drain stealing default for L2/HBM scenes; full GPU suite; bench lines
set -u
P=${1:-r2n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/${P}_pytest.log
B="timeout 600 python bench.py --warmup 3 --no-cpu-baseline"
for w in c2 c4 c5; do
$B --steps 8 --workload $w 2>&1 | tail -1 > gpurun_out/${P}_bench_${w}.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_${w}.json').read()); print('$w: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
$B --steps 32 --opt steal_smem=1 2>&1 | tail -1 > gpurun_out/${P}_bench_c2_steal_smem.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_c2_steal_smem.json').read()); print('c2 steal_smem: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
