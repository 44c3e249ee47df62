#!/usr/bin/env python
"""Development aid: top source lines of an `ncu --page source --csv --print-source cuda,sass` export by executed warp instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = None
out = []
for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != '':
        ie = hdr.index('Instructions Executed'); st = hdr.index('# Samples')
        try: out.append((int(r[ie]), int(r[st]), r[0], r[1].strip()[:150]))
        except ValueError: pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total warp instructions", tot, "samples", tots)
for o in sorted(out, reverse=True)[:top]:
    print("%9d %5.1f%% smp %5.1f%%  L%-5s %s" % (o[0], 100.0 * o[0] / max(tot, 1), 100.0 * o[1] / max(tots, 1), o[2], o[3]))
