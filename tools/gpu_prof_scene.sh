#!/usr/bin/env bash
# usage: tools/gpu_prof_scene.sh <tag> <scene> <res> <spp> <mb> <edges> [kernel-regex-for-full-capture ...]
# per-kernel table of ONE step (time, DRAM bytes, instructions, lanes/instr, issue-active) + optional full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; shift; SC=$1; RES=$2; SPP=$3; MB=$4; ED=$5; shift 5
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,gcc__average_cache_request_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum \
    --clock-control none --csv --log-file gpurun_out/${TAG}_step_kernels.csv python tools/one_step.py $SC $RES $SPP $MB $ED > gpurun_out/${TAG}_one_step.log 2>&1
python tools/summarize_step.py gpurun_out/${TAG}_step_kernels.csv gpurun_out/${TAG}_dram_traffic.json | tee gpurun_out/${TAG}_step_kernels.txt
python - "$TAG" <<'PY'
import csv, collections, sys
rows = list(csv.reader(open('gpurun_out/%s_step_kernels.csv' % sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mn, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
seen = set(); d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv and ('k_' in r[kn]): d.setdefault((r[idc], r[kn].split('(')[0]), {})[r[mn]] = r[mv]
for (i, k), m in d.items():
    if k in seen: continue
    seen.add(k)
    print(k, ' '.join('%s=%s' % (a.split('.')[0].replace('smsp__', '').replace('sm__', '').replace('gcc__average_cache_request_', 'gcc_'), b) for a, b in m.items() if 'dram' not in a and 'time' not in a))
PY
for K in "$@"; do
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/${TAG}_prof_$K -f \
      python tools/one_step.py $SC $RES $SPP $MB $ED > gpurun_out/${TAG}_ncu_full_$K.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$K.ncu-rep --page details --csv > gpurun_out/${TAG}_prof_${K}_details.csv 2>/dev/null
done
du -sh gpurun_out
