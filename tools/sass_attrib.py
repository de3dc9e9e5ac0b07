#!/usr/bin/env python
"""Attribute SASS instruction counts of one kernel to source lines through the inline chain.
usage: sass_attrib.py <nvdisasm --print-line-info-inline output> <kernel-substring> [depth]
Prints, for every call-site level, the inclusive instruction count per file:line (static code size,
the quantity that matters for the instruction-cache footprint of the megakernels)."""
import re, sys, collections
path, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fre = re.compile(r'//## File "([^"]+)", line (\d+)')
inkern = False
chain = []          # current group of File lines
pending = []
incl = collections.Counter(); leaf = collections.Counter(); total = 0
newgroup = True
for ln in open(path):
    if ln.startswith('.text.') or '.section' in ln:
        inkern = kern in ln
        continue
    if not inkern: continue
    m = fre.search(ln)
    if m:
        if newgroup: pending = []; newgroup = False
        pending.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', ln):
        if pending: chain = pending
        newgroup = True
        total += 1
        if chain:
            leaf[chain[0]] += 1
            for c in set(chain): incl[c] += 1
print('total instructions', total)
print('--- inclusive by file:line (call sites) ---')
for (f, l), c in incl.most_common(topn): print('%7d %5.1f%%  %s:%d' % (c, 100.0*c/total, f, l))
print('--- leaf by file ---')
byf = collections.Counter()
for (f, l), c in leaf.items(): byf[f] += c
for f, c in byf.most_common(20): print('%7d %5.1f%%  %s' % (c, 100.0*c/total, f))
