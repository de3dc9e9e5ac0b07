#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_full.log; tail -25 gpurun_out/pytest_full.log | cut -c1-300
timeout 300 python -m pytest tests/test_scene_build_gpu.py -m gpu -q -s 2>&1 | grep "records\|passed\|failed" | cut -c1-200
bash tools/gpu_ab.sh
echo "=== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c2.json | cut -c1-600
echo "=== bench c3"; timeout 900 python bench.py --steps 3 --warmup 3 --workload c3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c3.json | cut -c1-600
echo "=== bench c4"; timeout 900 python bench.py --steps 3 --warmup 3 --workload c4 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c4.json | cut -c1-600
