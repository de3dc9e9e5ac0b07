#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/persist_check.py 2>&1 | tail -20
for e in "" 1; do
  echo "RB_NO_PERSISTENT_PICK=$e"
  env ${e:+RB_NO_PERSISTENT_PICK=1} RB_EDGES=3 timeout 200 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -1 | cut -c1-150
  env ${e:+RB_NO_PERSISTENT_PICK=1} RB_EDGES=3 timeout 300 python tools/attrib.py teapot 512 32 2 2>&1 | tail -1 | cut -c1-150
  env ${e:+RB_NO_PERSISTENT_PICK=1} RB_EDGES=3 timeout 300 python tools/attrib.py bunny_box 512 16 5 2>&1 | tail -1 | cut -c1-150
  env ${e:+RB_NO_PERSISTENT_PICK=1} RB_EDGES=3 timeout 300 python tools/attrib.py glossy_room 256 16 2 2>&1 | tail -1 | cut -c1-150
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_full.log; tail -8 gpurun_out/pytest_full.log | cut -c1-300
