#!/usr/bin/env bash
# usage: tools/gpu_multi.sh N [workload]   -- bench.py on N GPUs of one box (tiles = headline, poses in the extra key)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; WL=${2:-c2}
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $WL 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -3 | tee gpurun_out/bench_ours_n${N}_${WL}.json | cut -c1-2500
