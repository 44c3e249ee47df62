#!/usr/bin/env python
"""Fingerprints of cell sets.

fingerprint(boxes)            count + sha256 over the sorted cell boxes (raw bits): order independent, used for the unions of
                              the deep-grid goldens (fingerprint_unions.py) where both sides enumerate the same merged cells.
leaf_fingerprint(indices)     count + sha256 over the sorted LEAF cells of a set, as integer lattice indices at the grid
                              accuracy: independent of how cells are merged, so a set computed in chunks can be compared.

As a program: bench_fingerprint.py <out.json> <dump> [<dump> ...] -- reads `ref_harness vanderpol worker ... dump=1` dumps
(the reference's UpperReachEvolveWorker::get_result() on chunks of the bench frontier), unions their reach / evolve sets leaf
by leaf and writes the fingerprint bench.py compares its own lap against."""
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


def leaf_fingerprint(idx):
    """idx: (n, dim) integer lattice indices of leaf cells (duplicates allowed)."""
    a = np.ascontiguousarray(np.asarray(idx, dtype=np.int64))
    a = a.reshape(len(idx), -1) if len(idx) else a.reshape(0, 0)
    if a.shape[0]:
        a = np.unique(a, axis=0)
    return {"cells": int(a.shape[0]), "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()}


def leaves_of_boxes(boxes, origin, lengths, acc):
    """Lattice indices of the accuracy-`acc` leaf cells covered by dyadic cell boxes (n, 2*dim)."""
    b = np.asarray(boxes, dtype=np.float64).reshape(len(boxes), -1)
    dim = b.shape[1] // 2
    w = np.asarray(lengths[:dim], dtype=np.float64) / float(1 << acc)
    o = np.asarray(origin[:dim], dtype=np.float64)
    lo = np.rint((b[:, 0::2] - o) / w).astype(np.int64)
    hi = np.rint((b[:, 1::2] - o) / w).astype(np.int64)          # exclusive
    out = []
    for k in range(b.shape[0]):
        axes = [np.arange(lo[k, i], hi[k, i], dtype=np.int64) for i in range(dim)]
        g = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, dim)
        out.append(g)
    return np.concatenate(out, axis=0) if out else np.zeros((0, dim), dtype=np.int64)


if __name__ == "__main__":
    out_path, dumps = sys.argv[1], sys.argv[2:]
    acc = None; grid = None; leaves = {"reach": [], "evolve": []}; cells = 0; sources = []
    for path in dumps:
        d = golden_io.load_dump(os.path.abspath(path))
        acc = d["config"]["acc"]; grid = d["grid"]["loc"]
        for which in ("reach", "evolve"):
            for loc in sorted(d["sets"].get(which, {})):
                leaves[which].append(leaves_of_boxes(d["sets"][which][loc], grid["origin"], grid["lengths"], acc))
        head = [l for l in open(path).read(4096).splitlines() if l.startswith("worker")]
        sources += head[:1]
        cells += int(head[0].split()[4]) if head else 0
    out = {}
    for which in ("reach", "evolve"):
        out[which] = leaf_fingerprint(np.concatenate(leaves[which], axis=0) if leaves[which] else np.zeros((0, 2), dtype=np.int64))
    out["frontier_cells"] = cells
    out["what"] = "leaf cells (lattice indices at accuracy %d) of the union over %d chunks of UpperReachEvolveWorker::get_result() of the compiled reference" % (acc, len(dumps))
    out["source"] = sources
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out))
