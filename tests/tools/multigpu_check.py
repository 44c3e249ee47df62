"""Multi-GPU check of the in-library sharded lap (ab200_comm_init + NCCL exchange of the cell keys on the device).
Launch with one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/tools/multigpu_check.py
Every rank runs the same analyser calls; the device laps are sharded over the ranks. Checked on every rank against the golden
vectors of the compiled reference: bouncing-ball outer_chain_reach (configs[1]), the van der Pol two-lap upper_reach_evolve and
the boost converter's upper reach (configs[4] system)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io  # noqa: E402
from conftest import load_pkg  # noqa: E402


def system(name):
    return open(os.path.join(golden_io.GOLDEN, name + ".sys")).read()


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_pkg()
    out = {"world": world}
    # configs[1]: bouncing ball chain reach, golden at accuracy 5 and 6
    for name, init in (("bouncingball_chain_acc5.txt", [0.0, 0.001, 1.999, 2.0]), ("bouncingball_chain_acc6.txt", [-2.0, 0.5, 0.5, 1.5])):
        d = golden_io.load_dump(name)
        an = pkg.HybridReachabilityAnalyser(system("bouncingball"), [-10.1, 10.1, -0.1, 2.1], d["config"]["acc"], device=local, settings={"arena_words": 65536})
        an.evolver.init_comm(dist)
        t0 = time.perf_counter()
        reach = an.outer_chain_reach({"freefall": init})
        t1 = time.perf_counter()
        _, rb = an.enumerate(reach)
        assert golden_io.boxes_equal_bits(rb, d["sets"]["reach"]["freefall"]), "rank %d: %s differs from the golden" % (rank, name)
        out[name] = {"seconds": t1 - t0, "cells_evolved": int(an.cells_evolved()), "reach_cells": int(rb.shape[0])}
        an.close()
    # van der Pol, accuracy 4 golden (one lap) and the 1,100-cell two-lap run at accuracy 6
    d = golden_io.load_dump("vanderpol_reach_acc4.txt")
    an = pkg.HybridReachabilityAnalyser(system("vanderpol"), [-3.0, 3.0, -3.0, 3.0], 4, device=local)
    an.evolver.init_comm(dist)
    reach, evolve = an.upper_reach_evolve({"loc": [-1.6, -1.3, 0.3, 0.8]}, d["config"]["T"], d["config"]["events"])
    assert golden_io.boxes_equal_bits(an.enumerate(reach)[1], d["sets"]["reach"]["loc"]) and golden_io.boxes_equal_bits(an.enumerate(evolve)[1], d["sets"]["evolve"]["loc"])
    an.close()
    an = pkg.HybridReachabilityAnalyser(system("vanderpol"), [-3.0, 3.0, -3.0, 3.0], 6, device=local, settings={"arena_words": 32768, "max_boxes": 320})
    an.evolver.init_comm(dist)
    an.upper_reach_evolve({"loc": [1.0, 1.001, 1.0, 1.001]}, 1e-3, 5)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    reach, evolve = an.upper_reach_evolve({"loc": [-1.0, 1.0, -1.0, 1.0]}, 2 * an.lock_to_grid_time, 5)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    n_reach = int(an.enumerate(reach)[1].shape[0])
    assert n_reach == 632, n_reach
    out["vanderpol acc 6, 2 laps"] = {"seconds": t1 - t0, "cells_evolved": int(an.cells_evolved()), "reach_cells": n_reach}
    an.close()
    # configs[4] system: boost converter upper reach, golden at accuracy 4
    d = golden_io.load_dump("boost_reach_acc4.txt")
    an = pkg.HybridReachabilityAnalyser(system("boost"), [0.0, 0.001, 0.0, 3.0, 0.0, 6.0], 4, device=local)
    an.evolver.init_comm(dist)
    reach, evolve = an.upper_reach_evolve({"incr,below_duty": [0.0, 0.0, 1.0, 1.0, 1.0, 1.0]}, d["config"]["T"], d["config"]["events"])
    rl, rb = an.enumerate(reach)
    for lname, ref in d["sets"]["reach"].items():
        assert golden_io.boxes_equal_bits(rb[rl == an.evolver.location(lname)], ref), "boost reach cells in %s" % lname
    out["boost upper reach acc 4"] = {"reach_cells": int(rb.shape[0])}
    an.close()
    dist.barrier()
    if rank == 0:
        out["verdict"] = "all ranks: cell sets identical to the goldens"
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
