#!/bin/bash
# round 2, visit l: no ticket atomics after exhaustion; stealing thresholds
set -u
P=${1:-r2l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "steal or async or large_scene" 2>&1 | tail -5 | tee gpurun_out/${P}_pytest.log
B="timeout 600 python bench.py --warmup 3 --no-cpu-baseline"
for s in 0 1 2 4 0 1 2 4; do
$B --steps 32 --opt steal=$s 2>&1 | tail -1 > gpurun_out/${P}_bench_steal$s.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_steal$s.json').read()); print('steal=$s: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
for w in c4 c5; do for s in 0 1 4; do
$B --steps 4 --workload $w --opt steal=$s 2>&1 | tail -1 > gpurun_out/${P}_bench_${w}_steal$s.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_${w}_steal$s.json').read()); print('$w steal=$s: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done; done
for s in 0 1 4; do
timeout 300 python tools/tail_probe.py profile=1 only=1080 steal=$s > gpurun_out/${P}_tail_probe_steal$s.txt 2>&1
grep "launch" gpurun_out/${P}_tail_probe_steal$s.txt | grep -v "end of"
done
