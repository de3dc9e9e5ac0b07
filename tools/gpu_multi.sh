#!/usr/bin/env bash
# Multi-GPU check (run with gpurun --gpus N): bench in both sharding modes + the reference arm under torchrun + the gpu tests.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== bench poses (weak scaling) N=$N"; timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-600
echo "=== bench tiles (one image over N GPUs) N=$N"; timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --mode tiles 2>&1 | tail -1 | tee gpurun_out/bench_ours_tiles_n$N.json | cut -c1-600
echo "=== bench N=1 on the same box"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_samebox.json | cut -c1-400
echo "=== reference arm under torchrun"; timeout 900 $TR --master-port 29543 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-600
echo "=== pytest -m gpu (multi-GPU tests included)"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
