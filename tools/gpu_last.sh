#!/bin/bash
set -u
timeout 200 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "async_kernel or cost_ordered or counters_and_determinism or octant_and_plain" 2>&1 | tail -4
python tools/tail_probe.py 2>&1 | cut -c1-150 | tee gpurun_out/s3c_tail_probe.txt
