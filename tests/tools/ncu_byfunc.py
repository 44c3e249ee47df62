#!/usr/bin/env python
"""Development aid: executed warp instructions per source function from an `ncu --page source --csv --print-source cuda,sass`
export (lines are matched by text against the current sources under release-1.0_b200/csrc)."""
import csv, sys, os, re, collections
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'release-1.0_b200', 'csrc')
text2loc = {}
funcs = {}
hdr_re = re.compile(r'^\s*(AB_HDN|AB_HD|__global__|inline|static)\b.*\b([A-Za-z_0-9]+)\s*\(')
for fn in sorted(os.listdir(root)):
    if not fn.endswith(('.cuh', '.cu', '.inc', '.h')): continue
    cur = '?'
    for i, l in enumerate(open(os.path.join(root, fn), errors='replace').read().split('\n')):
        m = hdr_re.match(l)
        if m and not l.strip().endswith(';'): cur = m.group(2)
        if l.startswith('#define AB_FILTER'): cur = l.split()[1].split('(')[0]
        text2loc.setdefault(l.strip()[:150], (fn, cur))
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; byf = collections.Counter(); bys = collections.Counter(); unk = 0
for r in rows:
    if len(r) > 8 and r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); st = hdr.index('# Samples'); continue
    if hdr is None or len(r) < 10 or r[0] == '': continue
    try: n = int(r[ie]); s = int(r[st])
    except ValueError: continue
    loc = text2loc.get(r[1].strip()[:150])
    if loc is None: unk += n; loc = ('?', r[1].strip()[:40])
    byf[loc] += n; bys[loc] += s
tot = sum(byf.values()); tots = sum(bys.values())
print('total', tot, 'unmatched', unk)
for k, v in byf.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print('%12d %5.1f%%  smp %5.1f%%  %s:%s' % (v, 100.0 * v / tot, 100.0 * bys[k] / max(tots, 1), k[0], k[1]))
