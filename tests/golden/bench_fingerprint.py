#!/usr/bin/env python
"""Fingerprint of a reach / evolve cell-set dump of oracle/harness/ref_harness.cpp (`worker ... dump=1`): the boxes of
the maximal dyadic cells, sorted, hashed as the raw bits of their endpoints. bench.py computes the same over the sets its
timed lap returned and compares."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import golden_io


def fingerprint(boxes):
    """boxes: (n, 2*dim) float64; order-independent (rows sorted), bit-exact."""
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64))
    b = b.reshape(len(boxes), -1) if len(boxes) else b.reshape(0, 0)
    if b.shape[0]:
        b = b[np.lexsort(b.T[::-1])]
    return {"cells": int(b.shape[0]), "sha256": hashlib.sha256(b.tobytes()).hexdigest()}


if __name__ == "__main__":
    d = golden_io.load_dump(os.path.abspath(sys.argv[1]))
    out = {}
    for which in ("reach", "evolve"):
        rows = [bx for loc in sorted(d["sets"].get(which, {})) for bx in d["sets"][which][loc]]
        out[which] = fingerprint(rows)
    head = open(sys.argv[1]).read(4096).splitlines()
    out["source"] = [l for l in head if l.startswith("worker")][:1]
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out))
