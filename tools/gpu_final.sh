#!/usr/bin/env bash
# round-end evidence on one box: GPU suite, per-kernel ncu tables of one step (C2 and teapot), one full capture each, bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_full.log; tail -4 gpurun_out/pytest_full.log | cut -c1-200
echo "=== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c2.json | cut -c1-400
echo "=== bench c3"; timeout 900 python bench.py --steps 3 --warmup 3 --workload c3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c3.json | cut -c1-300
echo "=== bench c4"; timeout 900 python bench.py --steps 3 --warmup 3 --workload c4 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c4.json | cut -c1-300
echo "=== bench reference arm (short)"; RB_REF_TOTAL_S=40 timeout 900 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_n1_c2.json | cut -c1-900
echo "=== ncu C2"; bash tools/gpu_prof_scene.sh r02_c2 shadow_blocker 512 64 1 3 k_primary_edge 2>&1 | grep -v "^at::\|^cub::\|^void at" | cut -c1-330 | head -40
echo "=== ncu teapot"; bash tools/gpu_prof_scene.sh r02_teapot teapot 256 32 2 3 k_bwd_sec_pick 2>&1 | grep -v "^at::\|^cub::\|^void at" | cut -c1-330 | head -40
