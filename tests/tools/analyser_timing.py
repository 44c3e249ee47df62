"""Analyser-level comparison (SURVEY 8d secondary figure): cells evolved per second of a whole multi-lap
upper_reach_evolve call, host orchestration included, for the library and for the compiled reference on the same box.
  python tests/tools/analyser_timing.py [acc] [laps]        (needs a B200 and oracle/_ref/ref_harness)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402

acc = int(sys.argv[1]) if len(sys.argv) > 1 else 5
laps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
init = [-1.0, 1.0, -1.0, 1.0]
pkg = load_pkg()
system = open(os.path.join(ROOT, "tests", "golden", "vanderpol.sys")).read()
an = pkg.HybridReachabilityAnalyser(system, [-3.0, 3.0, -3.0, 3.0], acc, settings={"arena_words": 32768, "max_boxes": 320, "hash_log2": 24})
T = laps * an.lock_to_grid_time
an.upper_reach_evolve({"loc": [1.0, 1.001, 1.0, 1.001]}, 1e-3, 5)           # warm-up (context, tuning upload)
t0 = time.perf_counter()
reach, evolve = an.upper_reach_evolve({"loc": init}, T, 5)
t1 = time.perf_counter()
ours = {"seconds": t1 - t0, "cells_evolved": an.cells_evolved(), "reach_cells": len(an.enumerate(reach)[1])}
ours["cells_per_s"] = ours["cells_evolved"] / ours["seconds"]
an.close()
harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
o = subprocess.check_output([harness, "vanderpol", "time", "acc=%d" % acc, "init=%s" % ",".join(str(x) for x in init), "laps=%d" % laps]).decode()
f = o.split()
ref = {"seconds": float(f[f.index("seconds") + 1]), "cells_evolved": int(f[f.index("cells_evolved") + 1]),
       "reach_cells": int(f[f.index("reach_cells") + 1]), "threads": int(f[f.index("threads") + 1])}
ref["cells_per_s"] = ref["cells_evolved"] / ref["seconds"]
print(json.dumps({"workload": "vanderpol upper_reach_evolve acc %d, %d laps, init %s" % (acc, laps, init), "ours": ours, "reference": ref,
                  "speedup": ours["cells_per_s"] / ref["cells_per_s"]}))
