"""epsilon_lower_chain_reach through the library beside the compiled reference's own run on the box's host cores (pruning off,
so that the reference's result does not depend on the order its worker threads finish): enclosures evolved per second of the
whole call, with the number of reach cells and the epsilon vector compared.
  python tests/tools/eps_lines.py [out.jsonl]        (needs a B200 and oracle/_ref/ref_harness)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
import golden_io  # noqa: E402

pkg = load_pkg()
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def system(name):
    return open(os.path.join(ROOT, "tests", "golden", name + ".sys")).read()


CASES = [
    ("vanderpol acc 3", "vanderpol", "vanderpol", None, [-3.0, 3.0, -3.0, 3.0], 3, {"loc": [1.0, 1.3, 1.0, 1.6]},
     ["acc=3", "prune=0", "init=1.0,1.3,1.0,1.6"]),
    ("watertank acc 3", "watertank", "watertank", "watertank_midparams", [0.0, 1.0, 4.5, 9.0], 3,
     {"flow,idle,rising": [1.0, 1.0, 6.0, 7.5], "flow,idle,falling": [0.0, 0.0, 6.0, 7.5]}, ["acc=3", "prune=0"]),
]

out_path = sys.argv[1] if len(sys.argv) > 1 else None
lines = []
for name, cfg, sysname, sub, domain, acc, init, args in CASES:
    an = pkg.HybridReachabilityAnalyser(system(sysname), domain, acc, settings={"arena_words": 262144, "max_boxes": 512})
    try:
        an.evolver.probe_fp64()                         # context warm-up
        t0 = time.perf_counter()
        reach, eps, laps = an.epsilon_lower_chain_reach(init, prune=False, substituted_system_text=system(sub) if sub else None)
        t1 = time.perf_counter()
        ours = {"seconds": t1 - t0, "enclosures_evolved": int(an.cells_evolved()), "laps": laps, "reach_cells": int(len(an.enumerate(reach)[1])),
                "evolve_ms": an.last_status.evolve_ms, "launches": int(an.last_status.launches)}
        ours["enclosures_per_s"] = ours["enclosures_evolved"] / ours["seconds"]
    finally:
        an.close()
    ref = {"error": "oracle/_ref/ref_harness missing"}
    if os.path.exists(HARNESS):
        tmp = "/tmp/eps_ref_%s.txt" % cfg
        t0 = time.perf_counter()
        subprocess.check_call([HARNESS, cfg, "epslower", "out=" + tmp] + args, timeout=900)
        t1 = time.perf_counter()
        d = golden_io.load_dump(tmp)
        first = open(tmp).read().split("threads")[1].split()[0]
        ref = {"seconds": d["config"]["ref_seconds"], "wall_seconds": t1 - t0, "threads": int(first),
               "reach_cells": sum(len(v) for v in d["sets"]["reach"].values())}
        ref["enclosures_per_s"] = ours["enclosures_evolved"] / ref["seconds"]
        # (library location indices follow the name order of the locations, as the reference's maps do)
        same_eps = all(golden_io.boxes_equal_bits(eps[i][:len(v)], v)
                       for i, (l, v) in enumerate(sorted(d["epsilon"].items())))
        ref["parity"] = "same reach cell count and epsilon bits" if (ref["reach_cells"] == ours["reach_cells"] and same_eps) else "MISMATCH"
    line = {"analysis": "epsilon_lower_chain_reach, pruning off", "config": name, "ours": ours, "reference": ref,
            "speedup": (ref["seconds"] / ours["seconds"]) if "seconds" in ref else None}
    print(json.dumps(line), flush=True)
    lines.append(line)
if out_path:
    with open(out_path, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
