#!/usr/bin/env python
"""ncu CSV of one step (tools/one_step.py) -> per-kernel table (launches, total time, DRAM bytes, instructions) on stdout
and profiles/<round>_dram_traffic.json (the `roofline.traffic` source of bench.py).
usage: summarize_step.py <ncu.csv> [out.json]"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mn, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split('(')[0]
    if 'cub' in name: name = 'cub::' + name.split('::')[-1].split('<')[0]
    name = name.replace('void ', '').replace('rb_lean::', '')[:48]  # (the lean instantiation shares the stage names)
    d = per.setdefault(name, collections.defaultdict(float))
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    d[r[mn]] += v
    if r[mn] == 'gpu__time_duration.sum': d['launches'] += 1
tot = sum(d['gpu__time_duration.sum'] for d in per.values())
traffic = {}
for k, d in per.items():
    traffic[k] = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
if len(sys.argv) > 2:
    json.dump({"source": "ncu --profile-from-start off, tools/one_step.py (one fwd+bwd step of C2), dram__bytes_read.sum + dram__bytes_write.sum summed over "
                         "the launches of each kernel", "dram_bytes_per_step": traffic}, open(sys.argv[2], 'w'), indent=1)
print('%-48s %5s %10s %6s %10s %10s %12s' % ('kernel', 'n', 'time ms', '%', 'dram rd MB', 'dram wr MB', 'warp instr'))
for k, d in sorted(per.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    t = d['gpu__time_duration.sum'] / 1e6
    rd, wr = d.get('dram__bytes_read.sum', 0), d.get('dram__bytes_write.sum', 0)
    print('%-48s %5d %10.3f %5.1f%% %10.1f %10.1f %12.4g' % (k, d['launches'], t, 100 * d['gpu__time_duration.sum'] / tot, rd / 1e6, wr / 1e6, d.get('smsp__inst_executed.sum', 0)))
print('total %.3f ms (serialised, cold-cache per-launch times: shares are meaningful, absolutes are not)' % (tot / 1e6))
