#!/bin/bash
# multi-GPU bench (N from $1), launched like the driver does: headline config (C2) and the 4K config (C3), fused peer reduce; parity of the reduced frame
set -u
N=${1:-2}
P=${2:-s3}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/${P}_gpus_n${N}.txt
echo "=== N=$N c2 reduce=peer"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 16 --warmup 3 --reduce peer --no-cpu-baseline 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/${P}_bench_n${N}_peer.json | cut -c1-250
echo "=== N=$N c3 (3840x2160) reduce=peer"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --workload c3 --steps 8 --warmup 3 --reduce peer --no-cpu-baseline 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/${P}_bench_c3_n${N}_peer.json | cut -c1-250
if [ "${NCCL:-0}" = "1" ]; then
echo "=== N=$N c2 reduce=nccl"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 16 --warmup 3 --reduce nccl --no-cpu-baseline 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/${P}_bench_n${N}_nccl.json | cut -c1-250
fi
echo "=== check_multi"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/check_multi.py 2>&1 | grep check_multi | tee gpurun_out/${P}_check_multi_n${N}.txt
