#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_full.log; cat gpurun_out/pytest_full.log
RB_EDGES=0,3 timeout 300 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -2 | cut -c1-260
RB_EDGES=3 timeout 600 python tools/attrib.py teapot 512 64 2 2>&1 | tail -1 | cut -c1-260
RB_EDGES=3 timeout 600 python tools/attrib.py bunny_box 512 32 5 2>&1 | tail -1 | cut -c1-260
RB_EDGES=3 timeout 600 python tools/attrib.py glossy_room 256 16 2 2>&1 | tail -1 | cut -c1-260
bash tools/gpu_prof_scene.sh r02b_teapot teapot 256 32 2 3 k_forward k_bwd_sec_pick | grep -v "^at::\|^cub::" | cut -c1-400
