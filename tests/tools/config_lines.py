"""One line per BASELINE.json config: cells evolved per second of the whole analyser call through the
library (host orchestration included) beside the compiled reference's own analyser on the box's host cores, with a parity
verdict (number of merged reach cells; the cell-by-cell comparisons live in tests/test_gpu_parity.py), plus the small-frontier
latency cases of VERDICT r1 (10-cell lap, 1,100-cell two-lap run).
  python tests/tools/config_lines.py [out.jsonl]        (needs a B200 and oracle/_ref/ref_harness)"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402

pkg = load_pkg()
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
REF_TIMEOUT = float(os.environ.get("AB200_REF_TIMEOUT", "150"))


def system(name):
    return open(os.path.join(ROOT, "tests", "golden", name + ".sys")).read()


def reference(cfg, args):
    if not os.path.exists(HARNESS):
        return {"error": "oracle/_ref/ref_harness missing"}
    try:
        o = subprocess.check_output([HARNESS, cfg, "time"] + args, timeout=REF_TIMEOUT).decode()
    except subprocess.TimeoutExpired:
        return {"error": "reference did not finish in %.0f s" % REF_TIMEOUT}
    f = o.split()
    r = {"seconds": float(f[f.index("seconds") + 1]), "cells_evolved": int(f[f.index("cells_evolved") + 1]),
         "reach_cells": int(f[f.index("reach_cells") + 1]), "threads": int(f[f.index("threads") + 1])}
    r["cells_per_s"] = r["cells_evolved"] / r["seconds"] if r["seconds"] > 0 else None
    return r


def ours(sysname, domain, acc, call, settings=None):
    an = pkg.HybridReachabilityAnalyser(system(sysname), domain, acc, settings=settings or {"arena_words": 65536, "max_boxes": 512})
    try:
        an.evolver.probe_fp64()                         # context warm-up
        t0 = time.perf_counter()
        reach = call(an)
        t1 = time.perf_counter()
        r = {"seconds": t1 - t0, "cells_evolved": int(an.cells_evolved()), "reach_cells": int(len(an.enumerate(reach)[1]))}
        r["cells_per_s"] = r["cells_evolved"] / r["seconds"]
        return r
    finally:
        an.close()


def line(name, cfgname, o, r):
    out = {"config": name, "ours": o, "reference": r}
    if "error" not in r:
        out["speedup"] = o["cells_per_s"] / r["cells_per_s"] if r.get("cells_per_s") else None
        out["parity"] = "same number of reach cells" if r["reach_cells"] == o["reach_cells"] else "MISMATCH %d vs %d" % (o["reach_cells"], r["reach_cells"])
    print(json.dumps(out), flush=True)
    return out


def main():
    lines = []
    # configs[0] watertank-hysteresis upper_reach as examples/watertank-hysteresis.cc sets it up, accuracy 3, horizon 2
    lines.append(line("configs[0] watertank upper_reach, accuracy 3, T=2 (the example's initial set)", "watertank",
                      ours("watertank", [0.0, 1.0, 4.5, 9.0], 3, lambda an: an.upper_reach_evolve({"flow,idle,rising": [1.0, 1.0, 6.0, 7.5], "flow,idle,falling": [0.0, 0.0, 6.0, 7.5]}, 2.0, 5)[0],
                           {"arena_words": 131072, "max_boxes": 512}),
                      reference("watertank", ["acc=3", "T=2"])))
    # configs[1] bouncing ball outer_chain_reach
    for acc, init in ((5, [0.0, 0.001, 1.999, 2.0]), (6, [-2.0, 0.5, 0.5, 1.5])):
        lines.append(line("configs[1] bouncingball outer_chain_reach, accuracy %d" % acc, "bouncingball",
                          ours("bouncingball", [-10.1, 10.1, -0.1, 2.1], acc, lambda an: an.outer_chain_reach({"freefall": init})),
                          reference("bouncingball", ["acc=%d" % acc] + (["init=%s" % ",".join(str(x) for x in init)] if acc == 6 else []))))
    # configs[2] at analyser level: two laps from [-1,1]^2 at accuracy 6 (1,100 cells evolved)
    lines.append(line("configs[2] vanderpol upper_reach_evolve, accuracy 6, 2 laps from [-1,1]^2", "vanderpol",
                      ours("vanderpol", [-3.0, 3.0, -3.0, 3.0], 6, lambda an: an.upper_reach_evolve({"loc": [-1.0, 1.0, -1.0, 1.0]}, 2 * an.lock_to_grid_time, 5)[0]),
                      reference("vanderpol", ["acc=6", "init=-1.0,1.0,-1.0,1.0", "laps=2"])))
    # configs[3] springs upper_reach from the example's point
    lines.append(line("configs[3] springs upper_reach, accuracy 3, T=40", "springs",
                      ours("springs", [0.0, 40.0, -2.0, 2.0, -3.0, 3.0, -1.0, 3.0, 0.0, 4.0], 3,
                           lambda an: an.upper_reach_evolve({"free1,free2": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0, 3.0]}, 40.0, 5)[0], {"arena_words": 262144, "max_boxes": 512}),
                      reference("springs", ["acc=3"])))
    # configs[4] boost converter: upper reach over 3 laps (its outer_chain_reach throws in the reference itself)
    lines.append(line("configs[4] boost converter upper_reach, accuracy 4, T=0.011", "boost",
                      ours("boost", [0.0, 0.001, 0.0, 3.0, 0.0, 6.0], 4, lambda an: an.upper_reach_evolve({"incr,below_duty": [0.0, 0.0, 1.0, 1.0, 1.0, 1.0]}, 0.011, 5)[0]),
                      reference("boost", ["acc=4", "chain=0", "T=0.011"])))
    # latency of a tiny frontier: 10 cells of the bench frontier (accuracy 7), one lap, through the worker seam
    an = pkg.HybridReachabilityAnalyser(system("vanderpol"), [-3.0, 3.0, -3.0, 3.0], 7, settings={"arena_words": 32768, "max_boxes": 320})
    try:
        initial = an.discretise({0: [-2.2, 2.2, -2.8, 2.8]})
        _, cells = an.enumerate(initial, mince=True)
        sel = np.linspace(0, cells.shape[0] - 1, 10).astype(np.int64)
        boxes = pkg.pack_boxes(np.ascontiguousarray(cells[sel]), 2)
        locs = np.zeros((10,), dtype=np.int32)
        an.upper_reach_evolve({0: [1.0, 1.001, 1.0, 1.001]}, 1e-3, 5)
        t0 = time.perf_counter()
        an.evolver.upper_reach_evolve(locs, boxes, an.lock_to_grid_time, 5)
        t1 = time.perf_counter()
        st = an.evolver.last_status
        out = {"config": "latency: 10 cells of the configs[2] frontier, one lap through ab200_upper_reach_evolve", "seconds": t1 - t0, "evolve_ms": st.evolve_ms, "steps": int(st.steps)}
        print(json.dumps(out), flush=True)
        lines.append(out)
    finally:
        an.close()
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
