#!/usr/bin/env bash
# round-2 second GPU call: full -m gpu suite without -x on the un-fused hit point, statistics of every stat case, first timings of C3 / C4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_full.log
for C in c2_all_vertices_secondary_stat glossy_room_secondary_stat c2_shadow_blocker_secondary_stat; do
  timeout 300 python tools/diag_stat.py $C gpurun_out/diag_${C}.npz 2>&1 | tail -5
done
RB_EDGES=0,3 timeout 300 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -2 | cut -c1-260
RB_EDGES=0,1,3 timeout 600 python tools/attrib.py teapot 512 64 2 2>&1 | tail -3 | cut -c1-260
RB_EDGES=0,1,3 timeout 600 python tools/attrib.py bunny_box 512 32 5 2>&1 | tail -3 | cut -c1-260
cat gpurun_out/pytest_full.log
