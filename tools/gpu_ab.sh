#!/usr/bin/env bash
# A/B of library variants on one box: stage times of C2, teapot, bunny box with both edge samplers
cd "$(dirname "$0")/.."
for l in "" redner_b200/_variants/*.so; do
  [ -z "$l" ] || [ -f "$l" ] || continue
  echo "LIB=${l:-main}"
  RB_LIB=$l RB_EDGES=3 timeout 200 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -1 | cut -c1-150
  RB_LIB=$l RB_EDGES=3 timeout 300 python tools/attrib.py teapot 512 32 2 2>&1 | tail -1 | cut -c1-150
  RB_LIB=$l RB_EDGES=3 timeout 300 python tools/attrib.py bunny_box 512 16 5 2>&1 | tail -1 | cut -c1-150
done
