#!/bin/bash
# round 2, multi-GPU visit: `gpurun --gpus N -- bash tools/gpu_r2_multi.sh N`: the multi-GPU tests (vn_multi_*, Renderer::SetDevices, the
# one-process-per-GPU peer reduce with epoch flags), then bench.py under torchrun at N ranks (weak line + strong-scaling frame + parity)
set -u
N=${1:-2}
P=${2:-r2m$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "=== pytest multi-gpu + grid"
timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s -k "multi or set_devices or one_process or grid_matches" 2>&1 | grep -v "^\s*$" | tail -40 | tee gpurun_out/${P}_pytest_multi.log
echo "=== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${P}_bench_n1.json | cut -c1-200
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/${P}_bench_n$N.json | cut -c1-600
python - <<PY
import json
for f in ("gpurun_out/${P}_bench_n1.json", "gpurun_out/${P}_bench_n$N.json"):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, "value %.0f e2e %.0f reduce_ms %.3f strong %s parity %s" % (d["value"], d["e2e"]["value"], d["config"]["reduce_ms"], d.get("strong_scaling", {}).get("value"), d.get("parity")))
    except Exception as e:
        print(f, "FAILED", e)
PY
ls -la gpurun_out/${P}_*
