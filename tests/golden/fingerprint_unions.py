#!/usr/bin/env python
"""Replaces the `ureach` / `ufinal` cell lists of a `ref_harness cells ... disc=2` dump by their fingerprints (count +
sha256 over the sorted boxes, bench_fingerprint.fingerprint), so that deep-grid goldens with 10^5 cells stay small.
Usage: fingerprint_unions.py <dump.txt> <out.txt>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_fingerprint import fingerprint


def main(src, dst):
    out = []; cur = None; rows = []

    def flush():
        if cur is not None:
            fp = fingerprint([[float.fromhex(x) for x in r] for r in rows])
            out.append("%s_fp loc %s count %d sha256 %s\n" % (cur[0], cur[1], fp["cells"], fp["sha256"]))

    for line in open(src):
        t = line.split()
        if t and t[0] in ("ureach", "ufinal"):
            flush(); cur = (t[0], t[2]); rows = []
        elif t and t[0] == "c" and cur is not None:
            rows.append(t[1:])
        else:
            flush(); cur = None; rows = []
            out.append(line)
    flush()
    open(dst, "w").writelines(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
