#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/tree_diff.py single_triangle shadow_blocker glossy_room teapot_geometry bunny_box_shifted hires_room 2>&1 | grep "^==\|record" | cut -c1-330 | head -40
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_full.log; tail -12 gpurun_out/pytest_full.log | cut -c1-300
