#!/usr/bin/env bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list + full capture of the dominant kernel.
# Keep gpurun_out/ below 64 MiB (one .ncu-rep with sources is ~33 MB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
echo "=== variants"
for l in "" redner_b200/_variants/lib_b3.so redner_b200/_variants/lib_b5.so; do
  [ -z "$l" ] || [ -f "$l" ] || continue
  echo "LIB=$l"
  RB_LIB=$l RB_EDGES=0,3 timeout 200 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -2 | cut -c1-125
  RB_LIB=$l RB_EDGES=0,3 timeout 200 python tools/attrib.py glossy_room 256 16 2 2>&1 | tail -2 | cut -c1-125
done
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== compare full size"; timeout 600 python tools/compare.py shadow_blocker --res 512 --spp 64 --edges 0 2>&1 | tail -5
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "=== ncu full k_backward"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward -s 1 -c 1 -o gpurun_out/prof_k_backward -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8; du -sh gpurun_out
