#!/bin/bash
# multi-GPU bench (N from $1), launched like the driver does: headline config (C2) and the 4K config (C3), fused peer reduce
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
echo "=== N=$N c2 reduce=peer"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 16 --warmup 3 --reduce peer 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/bench_s2_n${N}_peer.log | cut -c1-250
echo "=== N=$N c3 (3840x2160) reduce=peer"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --workload c3 --steps 8 --warmup 3 --reduce peer 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/bench_s2_c3_n${N}_peer.log | cut -c1-250
if [ "${NCCL:-0}" = "1" ]; then
echo "=== N=$N c2 reduce=nccl"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 16 --warmup 3 --reduce nccl 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/bench_s2_n${N}_nccl.log | cut -c1-250
fi
echo "=== check_multi"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/check_multi.py 2>&1 | grep check_multi | tee gpurun_out/check_multi_s2_n${N}.log
