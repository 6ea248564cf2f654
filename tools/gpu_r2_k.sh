#!/bin/bash
# round 2, visit k: sample stealing in the drain (A/B in one visit) + probe
set -u
P=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_config_parity.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/${P}_pytest.log
B="timeout 600 python bench.py --warmup 3 --no-cpu-baseline"
for s in 0 1 0 1; do
$B --steps 32 --opt steal=$s 2>&1 | tail -1 > gpurun_out/${P}_bench_steal$s.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_steal$s.json').read()); print('steal=$s: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
for w in c3 c4 c5; do for s in 0 1; do
$B --steps 4 --workload $w --opt steal=$s 2>&1 | tail -1 > gpurun_out/${P}_bench_${w}_steal$s.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_${w}_steal$s.json').read()); print('$w steal=$s: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done; done
timeout 300 python tools/tail_probe.py profile=1 only=1080 > gpurun_out/${P}_tail_probe.txt 2>&1
grep "launch" gpurun_out/${P}_tail_probe.txt
