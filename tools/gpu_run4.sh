#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_full.log; tail -5 gpurun_out/pytest_full.log
for l in "" redner_b200/_variants/*.so; do
  echo "LIB=${l:-main}"
  RB_LIB=$l RB_EDGES=3 timeout 200 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -1 | cut -c1-150
  RB_LIB=$l RB_EDGES=3 timeout 300 python tools/attrib.py teapot 512 32 2 2>&1 | tail -1 | cut -c1-150
  RB_LIB=$l RB_EDGES=3 timeout 300 python tools/attrib.py bunny_box 512 16 5 2>&1 | tail -1 | cut -c1-150
done
timeout 900 python tools/fd_check.py 2>&1 | tail -20
