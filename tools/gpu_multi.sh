#!/bin/bash
# multi-GPU bench (N from $1): launched like the driver does
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
for red in peer nccl; do
  echo "=== N=$N reduce=$red"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 16 --warmup 3 --reduce $red 2>&1 | grep -v -i warn | tail -4 | tee gpurun_out/bench_n${N}_${red}.log
done
echo "=== check_multi"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/check_multi.py 2>&1 | grep check_multi | tee gpurun_out/check_multi_n${N}.log
echo "=== N=$N reference arm"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 4 --warmup 1 2>&1 | grep -v -i warn | tail -2 | cut -c1-400
