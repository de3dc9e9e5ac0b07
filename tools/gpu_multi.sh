#!/usr/bin/env bash
# usage: tools/gpu_multi.sh N [workload ...]   -- bench.py on N GPUs of one box (tiles = headline, poses in the extra key)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
nvidia-smi -L | wc -l
for WL in "${@:-c2}"; do
  ST=10; [ "$WL" = c4 ] && ST=3; [ "$WL" = c5 ] && ST=3; [ "$WL" = c3 ] && ST=3
  echo "=== $WL N=$N"
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $ST --warmup 3 --workload $WL 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|NCCL version" | tail -2 | tee gpurun_out/bench_ours_n${N}_${WL}.json | cut -c1-1200
done
