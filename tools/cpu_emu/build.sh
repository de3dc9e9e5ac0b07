#!/usr/bin/env bash
# Host-compiled build of the per-sample device headers (see emu_shim.h): a development aid and the subject of
# tests/test_device_code_cpu.py.  RB_EMU_OUT = output path, RB_EMU_OPT = optimisation flags (default "-O2 -g").
set -e
cd "$(dirname "$0")"
DATA="$(cd ../../redner_b200/data && pwd)"
g++ ${RB_EMU_OPT:--O2 -g} -std=c++17 -fPIC -shared -w -include emu_shim.h -I/usr/local/cuda/include -I../../include \
    -DRB_DATA_DIR="\"$DATA\"" ${RB_EMU_FLAGS:-} emu.cpp -o ${RB_EMU_OUT:-libredner_b200_emu.so}
