"""Readers for the golden vectors written by oracle/harness/ref_harness.cpp (hex floats, bit-exact)."""
import gzip
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _open(name):
    path = name if os.path.isabs(name) else os.path.join(GOLDEN, name)
    if path.endswith(".gz"):
        return gzip.open(path, "rt")
    return open(path, "r")


def _f(tok):
    return float.fromhex(tok)


def load_dump(name):
    """Returns dict(config=..., grid={loc: {...}}, cells=[...], sets={'reach': {loc: boxes}, ...})."""
    out = {"config": {}, "grid": {}, "cells": [], "sets": {}}
    cur_cell = None
    cur_set = None
    with _open(name) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "config":
                out["config"]["name"] = t[1]
                for a, b in zip(t[2::2], t[3::2]):
                    out["config"][a] = _f(b) if a in ("T", "lock") else int(b)
            elif k == "grid":
                loc = t[2]; n = int(t[4]); p = 5
                g = {"n": n}
                assert t[p] == "origin"; g["origin"] = [_f(x) for x in t[p + 1:p + 1 + n]]; p += 1 + n
                assert t[p] == "lengths"; g["lengths"] = [_f(x) for x in t[p + 1:p + 1 + n]]; p += 1 + n
                assert t[p] == "hmad"; p += 1 + n
                assert t[p] == "maxstep"; g["maxstep"] = _f(t[p + 1]); p += 2
                assert t[p] == "refw"; g["refw"] = [_f(x) for x in t[p + 1:p + 1 + n]]
                out["grid"][loc] = g
            elif k == "frontier":
                out["config"]["frontier"] = int(t[1]); out["config"]["ignore_activations"] = int(t[3])
                for a, b in zip(t[4::2], t[5::2]):          # optional "lower L backward B"
                    out["config"][a] = int(b)
            elif k == "cell":
                cur_cell = {"id": int(t[1]), "loc": t[3], "box": [_f(x) for x in t[5:]], "reach": [], "final": [],
                            "reach_loc": [], "final_loc": [], "dreach": [], "dfinal": [], "error": None}
                out["cells"].append(cur_cell); cur_set = None
            elif k in ("r", "f"):
                key = "reach" if k == "r" else "final"
                cur_cell[key].append([_f(x) for x in t[2:]]); cur_cell[key + "_loc"].append(t[1])
            elif k == "x":
                cur_cell["error"] = " ".join(t[1:])
            elif k in ("dreach", "dfinal"):
                cur_set = cur_cell[k]
            elif k == "union":
                cur_cell = None
            elif k in ("ureach_fp", "ufinal_fp"):
                # fingerprint of a large union set (tests/golden/fingerprint_unions.py): count + sha256 of the sorted boxes
                out.setdefault("fp", {}).setdefault(k[:-3], {})[t[2]] = {"cells": int(t[4]), "sha256": t[6]}
                cur_set = None
            elif k == "cons|":
                out.setdefault("cons_text", []).append(line[len("cons| "):])
            elif k in ("reach", "evolve", "ureach", "ufinal", "initial"):
                cur_set = out["sets"].setdefault(k, {}).setdefault(t[2], [])
            elif k == "c" and cur_set is not None:
                cur_set.append([_f(x) for x in t[1:]])
            elif k == "epsilon":
                # "epsilon loc <name> v0 v1 ..": the HybridFloatVector of epsilon_lower_chain_reach
                out.setdefault("epsilon", {})[t[2]] = [_f(x) for x in t[3:]]
            elif k == "seconds":
                out["config"]["ref_seconds"] = float(t[1])
    return out


def boxes_equal_bits(a, b):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64)); b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
    if a.size == 0 and b.size == 0:
        return True
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
