#!/bin/bash
# round 2, visit j: tile-seed precompute + throughput-over-time probe
set -u
P=${1:-r2j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_config_parity.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${P}_pytest.log
timeout 300 python tools/tail_probe.py profile=1 only=1080 > gpurun_out/${P}_tail_probe.txt 2>&1
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${P}_bench.json
python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench.json').read()); print('bench: %.0f Mrays/s e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
