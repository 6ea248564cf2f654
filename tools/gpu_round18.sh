#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "pipelined or progressive or dropin or checkpoint" 2>&1 | tail -5
timeout 900 python bench.py --steps 64 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f ms/step %.3f cpu %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d.get('cpu_baseline')))"
timeout 900 python bench.py --steps 64 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
