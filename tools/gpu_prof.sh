#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attribution"; timeout 600 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -6
timeout 600 python tools/attrib.py glossy_room 256 16 2 2>&1 | tail -4
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "=== ncu full k_backward"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward -s 1 -c 1 -o gpurun_out/prof_k_backward -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail; du -sh gpurun_out
