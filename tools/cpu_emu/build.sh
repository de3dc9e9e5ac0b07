#!/usr/bin/env bash
# DEVELOPMENT AID ONLY: host-compiled emulator of the per-sample render logic (see emu_shim.h).
set -e
cd "$(dirname "$0")"
DATA="$(cd ../../redner_b200/data && pwd)"
g++ -O2 -g -std=c++17 -fPIC -shared -w -include emu_shim.h -I/usr/local/cuda/include -I../../include \
    -DRB_DATA_DIR="\"$DATA\"" ${RB_EMU_FLAGS:-} emu.cpp -o ${RB_EMU_OUT:-libredner_b200_emu.so}
