#!/usr/bin/env python
"""Join an ncu source-page CSV (per-SASS-instruction samples) with nvdisasm's inline line info to attribute
warp-stall samples and executed instructions of one kernel to call sites.
usage: ncu_attrib.py <report.ncu-rep> <lib.so> <kernel-substring> [file-to-group-by=rb_render.cuh]
Requires ncu, cuobjdump and nvdisasm on PATH (all in the CUDA toolkit)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1:4]
groupfile = sys.argv[4] if len(sys.argv) > 4 else 'rb_render.cuh'
outer = len(sys.argv) > 5 and sys.argv[5] == 'outer'  # group by the OUTERMOST frame in that file instead of the innermost
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if 'rb_kernels' in f][0]
dis = subprocess.run(['nvdisasm', '--print-line-info-inline', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
fre = re.compile(r'//## File "([^"]+)", line (\d+)')
are = re.compile(r'\s+/\*([0-9a-f]+)\*/\s+(\S.*?);')
chains = {}
inkern = False; pending = []; chain = []; newgroup = True
for ln in dis.splitlines():
    if ln.startswith('.text.') or '.section' in ln:
        inkern = kern in ln; continue
    if not inkern: continue
    m = fre.search(ln)
    if m:
        if newgroup: pending = []; newgroup = False
        pending.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m = are.match(ln)
    if m:
        if pending: chain = pending
        newgroup = True
        chains[int(m.group(1), 16)] = (chain, m.group(2))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Address' in r or '# Address' in r or any(c.strip() == 'Source' for c in r))
hdr = rows[hi]
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name: return i
    return None
c_addr = col('Address'); c_smp = col('Warp Stall Sampling (All Samples)') or col('# Samples') ; c_exec = col('Instructions Executed'); c_thr = col('Thread Instructions Executed')
c_noinst = col('stall_no_inst') 
stallcols = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_')]
tot = collections.Counter(); bysite = collections.defaultdict(collections.Counter); byleaf = collections.defaultdict(collections.Counter); stall = collections.Counter()
base = None
for r in rows[hi + 1:]:
    if len(r) <= c_addr: continue
    try: addr = int(r[c_addr], 16)
    except ValueError: continue
    if base is None: base = addr
    off = addr - base
    smp = float(r[c_smp] or 0) if c_smp is not None else 0
    ex = float(r[c_exec] or 0) if c_exec is not None else 0
    th = float(r[c_thr] or 0) if c_thr is not None else 0
    ch, txt = chains.get(off, ([], '?'))
    m_ = [('%s:%d' % c) for c in ch if c[0] == groupfile]
    site = (m_[-1] if outer else m_[0]) if m_ else 'other'
    leaf = ch[0][0] if ch else '?'
    for d, k in ((bysite, site), (byleaf, leaf)):
        d[k]['smp'] += smp; d[k]['ex'] += ex; d[k]['th'] += th; d[k]['n'] += 1
    tot['smp'] += smp; tot['ex'] += ex; tot['th'] += th
    for h, i in stallcols:
        try: stall[h] += float(r[i] or 0)
        except ValueError: pass
print('kernel %s: %d SASS rows, samples %.0f, warp-instr %.3g, lanes/instr %.1f' % (kern, len(chains), tot['smp'], tot['ex'], tot['th'] / max(tot['ex'], 1)))
print('--- stall reasons (samples)')
for h, v in stall.most_common(8): print('   %-28s %6.1f%%' % (h, 100 * v / max(sum(stall.values()), 1)))
for title, d in (('call site in ' + groupfile, bysite), ('leaf file', byleaf)):
    print('--- by %s:  samples%%  instr%%  lanes  static' % title)
    for k, v in sorted(d.items(), key=lambda kv: -kv[1]['smp'])[:25]:
        print('   %-26s %6.1f%% %6.1f%%  %5.1f  %6d' % (k, 100 * v['smp'] / max(tot['smp'], 1), 100 * v['ex'] / max(tot['ex'], 1), v['th'] / max(v['ex'], 1), v['n']))
