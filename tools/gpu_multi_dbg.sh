#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 16 --warmup 3 --reduce peer --no-cpu-baseline > gpurun_out/dbg_n${N}.log 2>&1
echo "rc=$?"; grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/dbg_n${N}.log | tail -3 | cut -c1-300
grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/dbg_n${N}.log | tail -1 > gpurun_out/s3_bench_n${N}_peer.json
