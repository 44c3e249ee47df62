"""Development aid: first differing reach box of a `cells` golden on the GPU, printed as hex floats."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
import golden_io
pkg = load_pkg()
system, name, domain = sys.argv[1], sys.argv[2], [float(x) for x in sys.argv[3].split(",")]
d = golden_io.load_dump(name)
an = pkg.HybridReachabilityAnalyser(open(os.path.join(golden_io.GOLDEN, system + ".sys")).read(), domain, d["config"]["acc"],
                                    settings={"arena_words": 262144, "max_boxes": 2048, "hash_log2": 22})
ev = an.evolver
for lname, g in d["grid"].items():
    ev.tune_settings(ev.location(lname), g["maxstep"], g["refw"], g["origin"], g["lengths"])
dim = ev.dims[0]
locs = [ev.location(c["loc"]) for c in d["cells"]]
boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
(roff, rbox, rloc), (foff, fbox, floc) = ev.reach_evolve(locs, boxes, d["config"]["T"], d["config"]["events"], d["config"]["ignore_activations"])
bad = 0
for i, c in enumerate(d["cells"]):
    mine = rbox[roff[i]:roff[i + 1], :2 * dim]; ref = np.asarray(c["reach"], dtype=np.float64).reshape(-1, 2 * dim)
    if golden_io.boxes_equal_bits(mine, ref):
        continue
    bad += 1
    if bad > 3:
        continue
    print("cell", c["id"], c["loc"], "boxes mine/ref", mine.shape[0], ref.shape[0])
    for k in range(min(mine.shape[0], ref.shape[0])):
        if not golden_io.boxes_equal_bits(mine[k], ref[k]):
            print(" first differing box", k, "loc", rloc[roff[i] + k], c["reach_loc"][k])
            print("  mine", " ".join(float(x).hex() for x in mine[k]))
            print("  ref ", " ".join(float(x).hex() for x in ref[k]))
            ulps = (mine[k].view(np.int64) - ref[k].view(np.int64))
            print("  ulp diff", list(ulps))
            break
print("cells", len(d["cells"]), "mismatching", bad)
