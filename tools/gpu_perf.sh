#!/usr/bin/env bash
# Perf-iteration session on the GPU box: A/B table over every library under redner_b200/_variants/, the GPU parity
# suite on the main build, the full-size forward compare, one short bench and (optional) ONE full ncu capture.
# usage: tools/gpu_perf.sh [kernel-regex-for-ncu|none]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREGEX=${1:-k_backward}
nvidia-smi -L
echo "=== variants (per-stage ms: fwd | bwd | primary edge | camera)"
for l in "" redner_b200/_variants/*.so; do
  [ -z "$l" ] || [ -f "$l" ] || continue
  echo "LIB=${l:-main}"
  RB_LIB=$l RB_EDGES=0,3 timeout 200 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -2 | cut -c1-140
  RB_LIB=$l RB_EDGES=3 timeout 200 python tools/attrib.py glossy_room 256 16 2 2>&1 | tail -1 | cut -c1-140
done
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== compare full size"; timeout 600 python tools/compare.py shadow_blocker --res 512 --spp 64 --edges 0 2>&1 | tail -5
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
if [ "$KREGEX" != none ]; then
  echo "=== ncu full $KREGEX"
  rm -f gpurun_out/*.ncu-rep
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -o gpurun_out/prof_$KREGEX -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
du -sh gpurun_out
