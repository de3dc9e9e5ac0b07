#!/usr/bin/env bash
# round-2 first GPU call: full -m gpu suite without -x, then the failing statistical case on library variants + emulator
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_full.log
C=c2_all_vertices_secondary_stat
timeout 300 python tools/diag_stat.py $C gpurun_out/diag_main.npz 2>&1 | tail -5
RB_NO_LEAN=1 timeout 300 python tools/diag_stat.py $C gpurun_out/diag_nolean.npz 2>&1 | tail -5
RB_BAND_BYTES=200000 timeout 300 python tools/diag_stat.py $C gpurun_out/diag_smallband.npz 2>&1 | tail -5
for v in nofma precise ieee; do
  timeout 300 python tools/diag_stat.py $C gpurun_out/diag_$v.npz --lib redner_b200/_variants/$v.so 2>&1 | tail -5
done
timeout 300 python tools/diag_stat.py $C gpurun_out/diag_f64.npz --lib redner_b200/libredner_b200_f64.so 2>&1 | tail -5
timeout 600 python tools/diag_stat.py $C gpurun_out/diag_emu.npz --emu tools/cpu_emu/libredner_b200_emu.so 2>&1 | tail -5
cat gpurun_out/pytest_full.log
