#!/usr/bin/env bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list + full capture of the dominant kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_ours.json
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
echo "=== ncu full k_backward"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward -s 1 -c 1 -o gpurun_out/prof_k_backward -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward -s 1 -c 1 -o gpurun_out/prof_k_forward -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_primary_edge -s 1 -c 1 -o gpurun_out/prof_k_primary_edge -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
