"""Device side of the randomised parity sweep: the cases tests/tools/parity_sweep.py generated (goldens of the compiled
reference, copied with a manifest into build/sweep/ -- built artefacts, they travel to the GPU box but are not committed) are
run through the library on the B200 and compared bit for bit: cell sets of every location and, for epsilon_lower_chain_reach,
the epsilon vector. A case the reference rejected must be rejected by the library.
  python tests/tools/device_sweep.py [seconds] [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
import golden_io  # noqa: E402

pkg = load_pkg()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
out_path = sys.argv[2] if len(sys.argv) > 2 else None
only = os.environ.get("SWEEP_ONLY")          # comma-separated golden basenames: run just these


def system(name):
    return open(os.path.join(ROOT, "tests", "golden", name + ".sys")).read()


def kv(args):
    return dict(a.split("=", 1) for a in args if "=" in a)


def boxof(s):
    return [float(x) for x in s.split(",")]


def compare(an, d, got, which):
    """got: CellSet; golden set `which` per location, bit for bit (an absent location must be empty)."""
    locs, boxes = an.enumerate(got)
    ref = d["sets"].get(which, {})
    total = 0
    for lname, cells in ref.items():
        mine = boxes[locs == an.evolver.location(lname)]
        if not golden_io.boxes_equal_bits(mine, cells):
            return "%s cells differ in %s (%d vs %d)" % (which, lname, mine.shape[0], len(cells))
        total += len(cells)
    if total != boxes.shape[0]:
        return "%s: %d cells vs %d in the golden" % (which, boxes.shape[0], total)
    return None


def run(case, arena_words=262144):
    k = case["kind"]; a = kv(case["args"]); acc = int(a["acc"])
    d = golden_io.load_dump(os.path.join(ROOT, case["golden"]))
    if k == "pend_upper":
        sysname, domain = "pendulum", [-3.0, 3.0, -3.0, 3.0]
        init = {"loc": boxof(a["init"])}
    elif k == "boost_upper":
        sysname, domain = "boost", [0.0, 0.001, 0.0, 3.0, 0.0, 6.0]
        init = {"incr,below_duty": boxof(a["init"])}
    elif k.startswith("vdp"):
        sysname, domain = "vanderpol", [-3.0, 3.0, -3.0, 3.0]
        init = {"loc": boxof(a["init"])} if "init" in a else None
    elif k.startswith("bb"):
        sysname, domain = "bouncingball", [-10.1, 10.1, -0.1, 2.1]
        init = {"freefall": boxof(a["init"])}
    elif k.startswith("sp"):
        sysname, domain = "springs", [0.0, 40.0, -2.0, 2.0, -3.0, 3.0, -1.0, 3.0, 0.0, 4.0]
        init = {"free1,free2": boxof(a["init"])}
    else:
        sysname, domain = "watertank", [0.0, 1.0, 4.5, 9.0]
        init = {"flow,idle,rising": boxof(a["initr"]), "flow,idle,falling": boxof(a["initf"])}
    cons = "".join(d.get("cons_text", []))
    an = pkg.HybridReachabilityAnalyser(system(sysname), domain, acc, settings={"arena_words": max(arena_words, 1048576) if k == "pend_upper" else arena_words, "max_boxes": 1024})
    try:
        try:
            if "lock" in a:
                an.set_lock_to_grid_time(float(a["lock"]))
            if k in ("vdp_upper", "wt_upper", "sp_upper", "pend_upper", "boost_upper"):
                r, e = an.upper_reach_evolve(init, float(a["T"]), 5)
                err = compare(an, d, r, "reach") or compare(an, d, e, "evolve")
            elif k == "vdp_cons_upper":
                r, e = an.upper_reach_evolve_from(an.discretise_constraint(cons), float(a["T"]), 5)
                err = compare(an, d, r, "reach") or compare(an, d, e, "evolve")
            elif k in ("vdp_lower", "bb_lower", "wt_lower"):
                r, e = an.lower_reach_evolve(init, float(a["T"]), 5)
                err = compare(an, d, r, "reach") or compare(an, d, e, "evolve")
            elif k == "vdp_cons_lower":
                r, e = an.lower_reach_evolve_constraint(cons, float(a["T"]), 5)
                err = compare(an, d, r, "reach") or compare(an, d, e, "evolve")
            elif k == "bb_chain":
                err = compare(an, d, an.outer_chain_reach(init), "reach")
            elif k in ("bb_bchain", "vdp_bupper"):
                err = compare(an, d, an.outer_chain_reach(init, backward=True), "reach")
            elif k in ("vdp_eps", "bb_eps", "vdp_cons_eps", "wt_eps", "sp_eps"):
                r, eps, laps = an.epsilon_lower_chain_reach(cons if k == "vdp_cons_eps" else init, prune=bool(int(a.get("prune", "1"))),
                                                            substituted_system_text=system("watertank_midparams") if k == "wt_eps" else None)
                err = compare(an, d, r, "reach")
                for lname, ref in d.get("epsilon", {}).items():
                    if not err and not golden_io.boxes_equal_bits(eps[an.evolver.location(lname)][:len(ref)], ref):
                        err = "epsilon differs in %s" % lname
            else:
                return "unknown kind"
        except pkg.Ab200Error as ex:
            if "ArenaExhausted" in str(ex) and arena_words < 2097152:
                # a device limit, reported as an error (coarse grids give large sets): once more with eight times the scratch
                an.close(); an = None
                case["arena_retry"] = arena_words * 8
                return run(case, arena_words * 8)
            return None if case["referr"] else "library error: %s" % ex
        if case["referr"]:
            return "the reference rejected this run, the library did not"
        return err
    finally:
        if an is not None:
            an.close()


cases = json.load(open(os.path.join(ROOT, "build", "sweep", os.environ.get("SWEEP_MANIFEST", "manifest.json"))))
# cheapest first within each kind, kinds interleaved, so that a short budget still covers every analysis
if only:
    cases = [c for c in cases if os.path.basename(c["golden"]) in only.split(",")]
cases.sort(key=lambda c: c["ref_seconds"])
t0 = time.time(); done = 0; bad = []; per_kind = {}
for c in cases:
    if time.time() - t0 > budget:
        break
    e = run(c)
    done += 1
    per_kind[c["kind"]] = per_kind.get(c["kind"], 0) + 1
    if e:
        bad.append({"case": c, "error": e}); print("MISMATCH", c["kind"], c["args"], e, flush=True)
summary = {"cases_run": done, "cases_total": len(cases), "mismatches": len(bad), "per_kind": per_kind, "arena_retries": [os.path.basename(c["golden"]) for c in cases[:done] if c.get("arena_retry")], "seconds": time.time() - t0,
           "reference_seconds_of_the_cases_run": sum(c["ref_seconds"] for c in cases[:done]), "bad": bad}
print(json.dumps(summary), flush=True)
if out_path:
    json.dump(summary, open(out_path, "w"))
sys.exit(1 if bad else 0)
