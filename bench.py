#!/usr/bin/env python
"""bench.py -- grid cells evolved / second for the upper_reach lap (BASELINE.json metric).

A "step" is one lock-to-grid lap of HybridReachabilityAnalyser::upper_reach over one frontier: every frontier cell is
evolved for lock_to_grid_time by the evolution kernel, its reach/final boxes are rasterised, de-duplicated and
sorted on the device. Workload = BASELINE.json configs[2]: van der Pol (mu=1), Taylor orders 6/6, >= 10^5 initial grid
cells (initial box [-2.2,2.2]x[-2.8,2.8] minced at accuracy 7 -> 101,520 cells), domain [-3,3]^2.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA)
  python bench.py --impl reference ...                     the reference's own CPU analyser (oracle/_ref), host cores

One JSON line is printed by rank 0. See DESIGN.md "Measurement" for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
_OUT = sys.stdout

METRIC = "grid cells evolved/sec (upper_reach)"
UNIT = "cells/s"
SYSTEM = "vanderpol"
DOMAIN = [-3.0, 3.0, -3.0, 3.0]
INIT = [-2.2, 2.2, -2.8, 2.8]
ACC = 7
EVENTS = 5
FRONTIER = 101520     # cells of the minced initial set at accuracy 7


def read_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""

    def __init__(self, index):
        threading.Thread.__init__(self)
        self.index = index; self.stop_flag = False; self.sm = []; self.max = 0; self.reasons = set()
        self.daemon = True

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], timeout=5).decode()
                p = [x.strip() for x in o.strip().split(",")]
                self.sm.append(float(p[0])); self.max = float(p[1])
                for n, v in zip(names, p[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max or None, "reasons": sorted(self.reasons)}


def frontier_cells(pkg, an, n_cells):
    """The minced initial frontier (uniform accuracy-ACC cells) exactly as the analyser enumerates it
    (outer_intersection_with + BDDTreeSet::mince), optionally sub-sampled in warp-sized blocks."""
    import numpy as np
    initial = an.discretise({0: INIT})
    _, cells = an.enumerate(initial, mince=True)
    an.free(initial)
    if n_cells and n_cells < cells.shape[0]:
        nb = cells.shape[0] // 32
        want = max(1, n_cells // 32)
        sel = np.linspace(0, nb - 1, want).astype(np.int64)
        idx = (sel[:, None] * 32 + np.arange(32)[None, :]).reshape(-1)
        cells = cells[idx]
    return np.ascontiguousarray(cells)


def fp64_sample(pkg, an, cells, locs_dtype):
    """Algorithmic FP64 operations per evolution step of THIS workload, measured in this run: a strided sample of the
    frontier goes through the warp build, whose kernels count the operations of the reference algorithm per primitive on
    the actual term counts (SURVEY.md 8d formulas; ab200_status.alg_flops). The lane build runs the same arithmetic and
    has no counters, so its FP64 rate is quoted with this per-step figure."""
    import numpy as np
    n = cells.shape[0]
    take = min(n, 16384)
    sel = np.linspace(0, n - 1, take).astype(np.int64)
    boxes = pkg.pack_boxes(np.ascontiguousarray(cells[sel]), 2)
    locs = np.zeros((take,), dtype=locs_dtype)
    old = os.environ.get("AB200_TUNE")
    os.environ["AB200_TUNE"] = "warp"
    try:
        an.evolver.upper_reach_evolve(locs, boxes, an.lock_to_grid_time, EVENTS)
        st = an.evolver.last_status
        rate = take / ((st.evolve_ms + st.raster_ms) * 1e-3) if st.evolve_ms > 0 else None
        return (st.alg_flops / st.steps) if st.steps else None, int(take), rate
    finally:
        if old is None:
            os.environ.pop("AB200_TUNE", None)
        else:
            os.environ["AB200_TUNE"] = old


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from conftest import load_pkg
    pkg = load_pkg()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this benchmark has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    system = open(os.path.join(ROOT, "tests", "golden", SYSTEM + ".sys")).read()
    an = pkg.HybridReachabilityAnalyser(system, DOMAIN, ACC, device=local, settings={"arena_words": args.arena_words, "max_boxes": args.max_boxes, "hash_log2": 25})
    ev = an.evolver
    all_cells = frontier_cells(pkg, an, args.cells)
    strong = args.scaling == "strong" and world > 1
    if world > 1:
        # One process per GPU; the library shards every lap: all ranks pass the SAME global frontier, the library deals blocks
        # of 64 cells round-robin to the ranks, each rank evolves and rasterises its share, and the sorted cell keys
        # are exchanged with ncclAllGather and de-duplicated on the device, so every rank returns the lap's full sets.
        ev.init_comm(dist)
    if strong:
        # STRONG scaling: ONE configs[2] frontier dealt to all GPUs.
        mine = all_cells
    elif world > 1:
        # WEAK scaling (the default; frontier cells are independent units): one configs[2] frontier PER GPU. The global
        # frontier of an N-GPU run is N translates of it (copy r shifted by r/N of a cell width along x, so the copies are
        # distinct enclosures with the same distribution of work); the library deals it, every GPU evolves 101,520 cells.
        w = all_cells[0, 1] - all_cells[0, 0]
        copies = []
        for r in range(world):
            c = all_cells.copy(); c[:, 0:2] += (float(r) / world) * w
            copies.append(c)
        mine = np.ascontiguousarray(np.concatenate(copies, axis=0))
    else:
        mine = all_cells
    n_mine = mine.shape[0]
    n_share = (n_mine + world - 1) // world                                 # cells this GPU evolves per lap
    boxes = pkg.pack_boxes(mine, 2)
    locs = np.zeros((n_mine,), dtype=np.int32)
    T = an.lock_to_grid_time * args.lap_scale
    an.lib.ab200_analyser_set_lock_time(an.handle, T)
    # one throw-away analyser call on a tiny set performs the tuning upload of a lap (tune_evolver)
    an.upper_reach_evolve({0: [1.0, 1.0 + 1e-3, 1.0, 1.0 + 1e-3]}, 1e-3, EVENTS)

    def lap():
        r, e = ev.upper_reach_evolve(locs, boxes, T, EVENTS)
        return r, e, ev.last_status

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        lap()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t0 = time.perf_counter()
    ev_ms = 0.0; ra_ms = 0.0; steps = 0; keys = 0; reach_n = 0; launches = 0; flops = 0.0; build = 0
    reach_sha = None
    for _ in range(args.steps):
        r, e, st = lap()
        ev_ms += st.evolve_ms; ra_ms += st.raster_ms; steps += st.steps; keys += st.keys_emitted; reach_n = len(r); launches += st.launches
        flops += st.alg_flops; build = st.build
        last_r, last_e = r, e
    barrier()
    t1 = time.perf_counter()
    sampler.stop_flag = True; sampler.join()
    wall = t1 - t0
    times = torch.tensor([wall, ev_ms, ra_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_mine) / world, float(steps), float(keys)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    wall, ev_ms, ra_ms = [float(x) for x in times.tolist()]
    n_total, steps_total, keys_total = [float(x) for x in tot.tolist()]
    e2e_value = n_total * args.steps / wall
    dev_value = n_total * args.steps / ((ev_ms + ra_ms) * 1e-3)
    # ---- parity of the benchmark's own output: fingerprint of the reach / evolve cell sets of the full configs[2] lap,
    # against the fingerprint of the reference's UpperReachEvolveWorker::get_result() on the same frontier
    # (tests/golden/vanderpol_bench_fingerprint.json, made by tests/golden/make_bench_fingerprint.sh)
    parity = None
    if world == 1 and args.lap_scale == 1.0 and not args.cells:
        parity = check_fingerprint(pkg, an, last_r, last_e)
    # ---- roofline of the dominant kernel (K_evolve): FP64 CUDA-core work. achieved = algorithmic FP64
    # operations of the reference algorithm / event time of the kernel; peak = separately rounded DMUL+DADD rate measured in
    # this run (interval arithmetic cannot fuse: every product and every sum carries its own rounding direction); the
    # DFMA rate is reported beside it.
    peaks = read_peaks()
    fp64 = fp64_fma = None
    try:
        fp64 = ev.probe_fp64(); fp64_fma = ev.probe_fp64(fma=True)
    except Exception:
        pass
    flops_per_step = None; flops_sample = None; warp_sample = None
    if flops > 0 and steps > 0:
        flops_per_step = flops / steps; flops_sample = "every cell of the timed laps (device counters)"
    else:
        fps, take, warp_rate = fp64_sample(pkg, an, all_cells, np.int32)
        flops_per_step = fps; flops_sample = "%d-cell strided sample of the same frontier through the warp build (device counters)" % take
        warp_sample = {"cells": take, "cells_per_s": warp_rate, "note": "the same frontier, strided, one untimed-for-the-headline lap through the warp-per-cell build"}
    alg_gflops = (steps / args.steps) * (flops_per_step or 0.0) / ((ev_ms / args.steps) * 1e-3) / 1e9 if ev_ms > 0 else 0.0
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "K_evolve_traffic.json")))
        if tj.get("frontier_cells_per_gpu") == int(n_share) and tj.get("build") == int(build) and args.lap_scale == 1.0:
            traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    build_name = {0: "warp (one warp per cell)", 1: "warp, two-word keys", 2: "lane (one lane per cell)"}.get(int(build), str(build))
    line = {
        "metric": METRIC, "value": dev_value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "vanderpol upper_reach lap, Taylor order 6/6, accuracy %d, %s (configs[2])" % (ACC, ("ONE frontier of %d cells dealt to %d GPUs" % (int(n_mine), world)) if strong else (("%d frontier cells per GPU" % int(n_share)) + ((", %d in all, dealt by the library" % int(n_mine)) if world > 1 else ""))),
                   "system": SYSTEM, "accuracy": ACC, "frontier_cells": int(n_total), "frontier_cells_per_gpu": int(n_share), "lap_time": T,
                   "evolution_build": build_name,
                   "l2_policy": "every lap rewrites its scratch arenas and box lists (GBs, far beyond L2); no flush needed"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_share * (80 + 4)), "d2h_bytes_per_step": int(8 * (reach_n + 0) + 24 * n_share)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "fp64", "achieved": alg_gflops, "peak": fp64, "unit": "GFLOP/s", "frac": (alg_gflops / fp64) if fp64 else None,
                     "traffic": traffic, "kernel": "K_evolve",
                     "peak_source": "measured in this run: DMUL.RD+DADD.RU chains (directed rounding cannot use FMA); DFMA chains give fp64_fma_gflops",
                     "fp64_fma_gflops": fp64_fma, "frac_of_fma_peak": (alg_gflops / fp64_fma) if fp64_fma else None,
                     "flops_per_step": flops_per_step, "flops_counted_on": flops_sample,
                     "evolve_ms_per_step": ev_ms / args.steps, "raster_sort_ms_per_step": ra_ms / args.steps,
                     "evolution_steps_per_s": steps_total / (ev_ms * 1e-3) if ev_ms > 0 else None,
                     "hbm": {"peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                             "raster_sort_keys_per_step": keys_total / args.steps if args.steps else None,
                             "raster_sort_gbs": (keys_total / args.steps) * 8.0 * (1 + 2 * 3 + 2) / ((ra_ms / args.steps) * 1e-3) / 1e9 if ra_ms > 0 else None}},
        "reach_cells": reach_n,
        "parity": parity,
    }
    if warp_sample:
        line["roofline"]["warp_build_sample"] = warp_sample
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.cpu_cells or 32 * (os.cpu_count() or 1))
        _OUT.write(json.dumps(line) + "\n"); _OUT.flush()
    an.close()
    if parity and parity.get("verdict") == "MISMATCH":
        raise SystemExit("bench.py: the lap's cell sets differ from the reference's fingerprint (see the parity field of the line)")
    if world > 1:
        dist.destroy_process_group()


def check_fingerprint(pkg, an, reach, evolve):
    """Fingerprint (leaf-cell count + sha256 over the sorted lattice indices) of the reach / evolve cell sets of the lap,
    against the committed fingerprint of the reference's own UpperReachEvolveWorker::get_result() on the same 101,520-cell
    frontier (tests/golden/vanderpol_bench_fingerprint.json, made by tests/golden/make_bench_fingerprint.sh from the
    compiled reference)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from bench_fingerprint import leaf_fingerprint
    path = os.path.join(ROOT, "tests", "golden", "vanderpol_bench_fingerprint.json")
    try:
        want = json.load(open(path))
    except Exception:
        return {"verdict": "unchecked", "why": "no committed fingerprint"}
    if want.get("frontier_cells") != FRONTIER:
        return {"verdict": "unchecked", "why": "the committed fingerprint covers %s of the %d frontier cells" % (want.get("frontier_cells"), FRONTIER)}
    out = {}
    for name, cs in (("reach", reach), ("evolve", evolve)):
        out[name] = leaf_fingerprint(an.evolver.decode(cs.keys)[:, 1:3])
    ok = all(out[k]["sha256"] == want[k]["sha256"] and out[k]["cells"] == want[k]["cells"] for k in ("reach", "evolve"))
    return {"verdict": "identical to the reference" if ok else "MISMATCH", "reach": out["reach"], "evolve": out["evolve"],
            "reference": {k: want[k] for k in ("reach", "evolve")}}


def cpu_baseline(n_cells, threads=None):
    """The reference's own CPU implementation at the same seam as ab200_upper_reach_evolve:
    UpperReachEvolveWorker::get_result() (include/workers.h:51-81) of the compiled, unmodified reference
    (oracle/_ref/ref_harness worker), with all host threads, on a strided sample of the same minced frontier."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/ref_harness missing"}
    cores = threads or os.cpu_count() or 1
    env = dict(os.environ); env["ORACLE_THREADS"] = str(cores)
    total_frontier = FRONTIER
    stride = max(1, total_frontier // n_cells)
    cmd = [harness, SYSTEM, "worker", "acc=%d" % ACC, "ncells=%d" % n_cells, "stride=%d" % stride, "init=%s" % ",".join(str(x) for x in INIT)]
    o = subprocess.check_output(cmd, env=env).decode()
    f = [l for l in o.splitlines() if l.startswith("worker")][0].split()
    secs = float(f[2]); done = int(f[4]); thr = int(f[-1])
    return {"value": done / secs, "unit": UNIT, "cores": thr, "kind": "reference",
            "sample": "UpperReachEvolveWorker::get_result() of the compiled reference on %d of the %s frontier cells (stride %d), one lap, %d threads, %.1f s" % (done, f[6], stride, thr, secs)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # Sample size: as many cells as keep the whole --steps/--warmup run within about three minutes, at most 2,048 per step.
    # The rate is probed on 8 cells per host thread first; it FALLS with the sample size, because the worker discretises and
    # adjoins every cell's result under one mutex (include/workers.h:141-150) into a set that keeps growing: 60 cells/s on 512
    # cells, 26 cells/s on 7,399 cells with 16 threads (gpurun_out/r2_bench_reference.json), so small samples flatter the
    # reference.
    cores = os.cpu_count() or 1
    probe = cpu_baseline(8 * cores)
    rate = probe["value"] or 50.0
    runs = args.steps + min(args.warmup, 1)
    n = args.cpu_cells or int(max(16 * cores, min(2048, 120.0 * rate / max(runs, 1))))
    vals = []
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(n)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(n)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "vanderpol upper_reach lap, Taylor order 6/6, accuracy %d, bounded sample (%d cells per step) of the %d-cell frontier (configs[2])" % (ACC, n, FRONTIER),
                       "system": SYSTEM, "accuracy": ACC},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _OUT.write(json.dumps(line) + "\n"); _OUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells", type=int, default=0, help="subsample the frontier (0 = the full 101,520 cells)")
    ap.add_argument("--cpu-cells", type=int, default=0, help="CPU sample size (0 = 3 cells per host thread)")
    ap.add_argument("--arena-words", type=int, default=32768)
    ap.add_argument("--max-boxes", type=int, default=320)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N>1: one frontier per GPU (weak, default) or one frontier dealt to all GPUs (strong)")
    ap.add_argument("--lap-scale", type=float, default=1.0, help="PROFILING ONLY: shorten the lap so that ncu can replay the kernel; not a bench value")
    args = ap.parse_args()
    # Rank 0 prints ONE JSON line on stdout and nothing else: whatever libraries write to file descriptor 1 while the run lasts
    # (NCCL prints its version there when NCCL_DEBUG is set) is sent to stderr; the line goes to the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _OUT
    _OUT = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
