#!/usr/bin/env bash
# Perf-iteration session on the GPU box: A/B table over every library under redner_b200/_variants/, the GPU parity
# suite on the main build, the full-size forward compare, one short bench and (optional) ONE full ncu capture.
# usage: tools/gpu_perf.sh [kernel-regex-for-ncu|none]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREGEX=${1:-k_primary_edge}
nvidia-smi -L
echo "=== variants (per-stage ms: fwd | bwd | primary edge | camera)"
for l in "" redner_b200/_variants/*.so; do
  [ -z "$l" ] || [ -f "$l" ] || continue
  echo "LIB=${l:-main}"
  RB_LIB=$l RB_EDGES=0,3 timeout 200 python tools/attrib.py shadow_blocker 512 64 1 2>&1 | tail -2 | cut -c1-170
  RB_LIB=$l RB_EDGES=3 timeout 200 python tools/attrib.py glossy_room 256 16 2 2>&1 | tail -1 | cut -c1-170
  RB_LIB=$l RB_EDGES=3 timeout 300 python tools/attrib.py hires_room 512 16 2 2>&1 | tail -1 | cut -c1-220
done
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== compare full size"; timeout 600 python tools/compare.py shadow_blocker --res 512 --spp 64 --edges 0 2>&1 | tail -5
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
echo "=== one profiled step: per-kernel time, DRAM bytes, instruction supply"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,gcc__average_cache_request_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/step_kernels.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
python tools/summarize_step.py gpurun_out/step_kernels.csv gpurun_out/dram_traffic.json
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/step_kernels.csv')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mn, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
seen = set()
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv and r[kn].startswith('k_'): d.setdefault((r[idc], r[kn].split('(')[0]), {})[r[mn]] = r[mv]
for (i, k), m in d.items():
    if k in seen: continue
    seen.add(k)
    print(k, ' '.join('%s=%s' % (a.split('.')[0].replace('smsp__', '').replace('sm__', '').replace('gcc__average_cache_request_', 'gcc_'), b) for a, b in m.items() if 'dram' not in a and 'time' not in a))
PY
if [ "$KREGEX" != none ]; then
  echo "=== ncu full $KREGEX"
  rm -f gpurun_out/*.ncu-rep
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -o gpurun_out/prof_$KREGEX -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
du -sh gpurun_out
echo "=== e2e breakdown"; timeout 600 python tools/e2e_breakdown.py 2>&1 | tail -45
