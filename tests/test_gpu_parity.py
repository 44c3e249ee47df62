"""Parity of the CUDA path (through the C ABI) against golden vectors produced by the compiled reference.
Integer / index work (cell sets) and every box endpoint must be bit-exact."""
import os

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def _system(name="vanderpol"):
    return open(os.path.join(golden_io.GOLDEN, name + ".sys")).read()


def _evolver(pkg, dump, system="vanderpol", **settings):
    ev = pkg.Evolver(_system(system), settings=settings or None)
    for name, g in dump["grid"].items():
        ev.tune_settings(ev.location(name), g["maxstep"], g["refw"], g["origin"], g["lengths"])
    ev.set_accuracy(dump["config"]["acc"])
    return ev


@pytest.mark.parametrize("system,name,arena", [("vanderpol", "vanderpol_cells_acc6.txt.gz", 16384), ("vanderpol", "vanderpol_cells_acc3.txt", 131072),
                                               ("vanderpol", "vanderpol_cells_acc7.txt.gz", 12288),
                                               ("bouncingball", "bouncingball_cells_acc4.txt", 16384), ("bouncingball", "bouncingball_cells_acc5.txt", 16384),
                                               ("pendulum", "pendulum_cells_acc4.txt", 1048576)])
def test_reach_evolve_boxes_bit_exact(pkg, system, name, arena):
    """Per-enclosure reach / final boxes. The bouncing-ball vectors run in chain-reach mode (ignore_activations) and
    exercise the urgent-guard path: activity tests, touching-time bisection and implicit crossing-time models."""
    d = golden_io.load_dump(name)
    ev = _evolver(pkg, d, system=system, arena_words=arena, max_boxes=320)
    try:
        dim = ev.dims[0]
        locs = [ev.location(c["loc"]) for c in d["cells"]]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
        events = 1 if d["config"]["chain"] else d["config"]["events"]
        (roff, rbox, rloc), (foff, fbox, floc) = ev.reach_evolve(locs, boxes, d["config"]["T"], events, d["config"]["ignore_activations"])
        for i, c in enumerate(d["cells"]):
            mine_r = rbox[roff[i]:roff[i + 1], :2 * dim]
            mine_f = fbox[foff[i]:foff[i + 1], :2 * dim]
            assert golden_io.boxes_equal_bits(mine_r, c["reach"]), "reach boxes of cell %d" % c["id"]
            assert golden_io.boxes_equal_bits(mine_f, c["final"]), "final boxes of cell %d" % c["id"]
        assert ev.last_status.steps == sum(len(c["reach"]) for c in d["cells"])
    finally:
        ev.close()


@pytest.mark.parametrize("system,name", [("bouncingball", "bouncingball_cells_acc4_lower.txt"),
                                         ("bouncingball", "bouncingball_cells_acc4_lower_events.txt"),
                                         ("watertank", "watertank_cells_acc3_lower.txt"),
                                         ("vanderpol", "vanderpol_cells_acc4_backward.txt"),
                                         ("bouncingball", "bouncingball_cells_acc4_backward.txt")])
def test_reach_evolve_lower_semantics_and_backward_bit_exact(pkg, system, name):
    """EvolverInterface::reach_evolve(..., LOWER_SEMANTICS): per-enclosure reach / final boxes, location by location (the
    reference's list sets are grouped by location, within a location in insertion order)."""
    d = golden_io.load_dump(name)
    ev = _evolver(pkg, d, system=system, arena_words=131072, max_boxes=320, semantics=d["config"].get("lower", 0))
    try:
        dim = ev.dims[0]
        locs = [ev.location(c["loc"]) for c in d["cells"]]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
        (roff, rbox, rloc), (foff, fbox, floc) = ev.reach_evolve(locs, boxes, d["config"]["T"], d["config"]["events"],
                                                                  ignore_activations=bool(d["config"]["ignore_activations"]),
                                                                  backward=bool(d["config"].get("backward", 0)))
        for i, c in enumerate(d["cells"]):
            for mine_b, mine_l, ref_b, ref_l, what in ((rbox[roff[i]:roff[i + 1], :2 * dim], rloc[roff[i]:roff[i + 1]], c["reach"], c["reach_loc"], "reach"),
                                                       (fbox[foff[i]:foff[i + 1], :2 * dim], floc[foff[i]:foff[i + 1]], c["final"], c["final_loc"], "final")):
                ref_l = np.array([ev.location(x) for x in ref_l], dtype=np.int32)
                ref_b = np.array(ref_b, dtype=np.float64).reshape(-1, 2 * dim)
                assert len(mine_l) == len(ref_l), "%s count of cell %d" % (what, c["id"])
                for l in range(ev.n_locations):
                    assert golden_io.boxes_equal_bits(mine_b[mine_l == l], ref_b[ref_l == l]), "%s boxes of cell %d in location %d" % (what, c["id"], l)
    finally:
        ev.close()


def _leaf_set(an, cellset):
    return set(int(k) for k in cellset.keys)


@pytest.mark.parametrize("name,init,laps", [("vanderpol_reach_acc3.txt", [1.0, 1.3, 1.0, 1.6], 2),
                                              ("vanderpol_reach_acc4.txt", [-1.6, -1.3, 0.3, 0.8], 1)])
def test_analyser_upper_reach_evolve_identical_cells(pkg, name, init, laps):
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system(), [-3.0, 3.0, -3.0, 3.0], d["config"]["acc"], settings={"arena_words": 131072})
    try:
        assert an.lock_to_grid_time == d["config"]["lock"]
        g = d["grid"]["loc"]
        assert list(an.grid_lengths[0][:2]) == g["lengths"] and list(an.grid_origin[0][:2]) == g["origin"]
        reach, evolve = an.upper_reach_evolve({"loc": init}, d["config"]["T"], d["config"]["events"])
        _, rb = an.enumerate(reach)
        _, eb = an.enumerate(evolve)
        assert golden_io.boxes_equal_bits(rb, d["sets"]["reach"]["loc"])
        assert golden_io.boxes_equal_bits(eb, d["sets"]["evolve"]["loc"])
    finally:
        an.close()


EPS_CASES = [
    ("bouncingball", "bouncingball_epslower_acc4.txt", [-10.1, 10.1, -0.1, 2.1], {"freefall": [0.0, 0.001, 1.999, 2.0]}, True, None),
    ("springs", "springs_epslower_acc2_lock1.txt", [0.0, 40.0, -2.0, 2.0, -3.0, 3.0, -1.0, 3.0, 0.0, 4.0],
     {"free1,free2": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0, 3.0]}, True, None),      # two-word keys; lock time 1 (3 laps)
    ("vanderpol", "vanderpol_epslower_acc3.txt", [-3.0, 3.0, -3.0, 3.0], {"loc": [1.0, 1.3, 1.0, 1.6]}, False, None),
    ("watertank", "watertank_epslower_acc3.txt", [0.0, 1.0, 4.5, 9.0],
     {"flow,idle,rising": [1.0, 1.0, 6.0, 7.5], "flow,idle,falling": [0.0, 0.0, 6.0, 7.5]}, False, "watertank_midparams"),
    ("watertank", "watertank_epslower_acc3_prune.txt", [0.0, 1.0, 4.5, 9.0],
     {"flow,idle,rising": [1.0, 1.0, 6.0, 7.5], "flow,idle,falling": [0.0, 0.0, 6.0, 7.5]}, True, "watertank_midparams"),
]


@pytest.mark.parametrize("system,name,domain,init,prune,substituted", EPS_CASES)
def test_epsilon_lower_chain_reach_identical_cells_and_epsilon(pkg, system, name, domain, init, prune, substituted):
    """HybridReachabilityAnalyser::epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627) with
    LowerReachEpsilonWorker (include/workers.h:171-329) on the device: the final TaylorSets of a lap are handed to the next lap
    as enclosure records. The reached cells and the epsilon vector equal the reference's bit for bit; with pruning on, the
    kept enclosures follow glibc's rand() sequence as in a fresh single-worker reference process."""
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system(system), domain, d["config"]["acc"], settings={"arena_words": 262144, "max_boxes": 512})
    try:
        if an.lock_to_grid_time != d["config"]["lock"]:
            an.set_lock_to_grid_time(d["config"]["lock"])          # (a golden taken with lock=...)
        reach, eps, laps = an.epsilon_lower_chain_reach(init, prune=prune, substituted_system_text=_system(substituted) if substituted else None)
        rl, rb = an.enumerate(reach)
        assert laps >= 2 and an.last_status.launches >= laps
        for lname, ref in d["sets"]["reach"].items():
            mine = rb[rl == an.evolver.location(lname)]
            assert golden_io.boxes_equal_bits(mine, ref), "reach cells of %s in %s" % (system, lname)
        assert sum(len(v) for v in d["sets"]["reach"].values()) == rb.shape[0]
        for lname, ref in d["epsilon"].items():
            mine = eps[an.evolver.location(lname)][:len(ref)]
            assert golden_io.boxes_equal_bits(mine, ref), "epsilon of %s in %s: %r vs %r" % (system, lname, mine, ref)
        assert an.evolver.settings.semantics == 0
    finally:
        an.close()


def test_epsilon_lower_chain_reach_domain_error_as_the_reference(pkg):
    """The boost converter's lower reach leaves the widened domain in the first lap: the reference throws
    (src/reachability_analyser.cpp:596-599, golden boost_epslower_acc3_domain_error.txt), the library returns AB200_E_DOMAIN."""
    g = open(os.path.join(golden_io.GOLDEN, "boost_epslower_acc3_domain_error.txt")).read()
    assert "The lower reach is not inside the over-approximated domain" in g
    an = pkg.HybridReachabilityAnalyser(_system("boost"), [0.0, 0.001, 0.0, 3.0, 0.0, 6.0], 3, settings={"arena_words": 262144, "max_boxes": 512})
    try:
        with pytest.raises(pkg.Ab200Error) as ei:
            an.epsilon_lower_chain_reach({"incr,below_duty": [0.0, 0.0, 1.0, 1.0, 1.0, 1.0]}, prune=True)
        assert ei.value.code == pkg.AB200_E_DOMAIN and "lower reach is not inside" in str(ei.value)
    finally:
        an.close()


def test_analyser_raises_the_scratch_limit_instead_of_failing(pkg):
    """The reference has no per-cell scratch limit. An analyser lap that ends with AB200_E_ARENA is run again with eight
    times the arena (with_arena_growth, ab200_analyser.cu) and gives the golden's cells; the evolver-level entry points keep
    the strict error (test_error_paths)."""
    d = golden_io.load_dump("vanderpol_reach_acc4.txt")
    an = pkg.HybridReachabilityAnalyser(_system("vanderpol"), [-3.0, 3.0, -3.0, 3.0], 4, settings={"arena_words": 1024, "max_boxes": 512})
    try:
        reach, evolve = an.upper_reach_evolve({"loc": [-1.6, -1.3, 0.3, 0.8]}, d["config"]["T"], d["config"]["events"])
        assert golden_io.boxes_equal_bits(an.enumerate(reach)[1], d["sets"]["reach"]["loc"])
        assert golden_io.boxes_equal_bits(an.enumerate(evolve)[1], d["sets"]["evolve"]["loc"])
        assert an.evolver.get_settings().arena_words >= 8192
    finally:
        an.close()


def test_lower_analyses_from_a_constraint_set(pkg):
    """lower_reach_evolve and epsilon_lower_chain_reach from a HybridBoundedConstraintSet: the midpoints of the initial cells
    that the set definitely covers (src/reachability_analyser.cpp:1218-1243). The set lies in the default (large) initial box,
    whose paving has its finest levels merged, so BDDTreeSet::outer_restrict stops above the accuracy (BDDTreeSet::depth(),
    src/bdd_set.cpp:1171-1190): the initial cells are coarser than the grid's and only 5 of 12 midpoints are covered."""
    d = golden_io.load_dump("vanderpol_cons_disc_lower_reach_acc4.txt")
    an = pkg.HybridReachabilityAnalyser(_system("vanderpol"), [-3.0, 3.0, -3.0, 3.0], 4, settings={"arena_words": 262144, "max_boxes": 512})
    try:
        text = "".join(d["cons_text"])
        il, ib = an.enumerate(an.discretise_constraint(text))
        assert golden_io.boxes_equal_bits(ib, d["sets"]["initial"]["loc"])
        reach, evolve = an.lower_reach_evolve_constraint(text, d["config"]["T"], d["config"]["events"])
        assert an.cells_evolved() == 5
        assert golden_io.boxes_equal_bits(an.enumerate(reach)[1], d["sets"]["reach"]["loc"])
        assert golden_io.boxes_equal_bits(an.enumerate(evolve)[1], d["sets"]["evolve"]["loc"])
        e = golden_io.load_dump("vanderpol_cons_disc_epslower_acc4.txt")
        reach, eps, laps = an.epsilon_lower_chain_reach("".join(e["cons_text"]), prune=True)
        assert laps == 3
        assert golden_io.boxes_equal_bits(an.enumerate(reach)[1], e["sets"]["reach"]["loc"])
        assert golden_io.boxes_equal_bits(eps[0][:2], e["epsilon"]["loc"])
    finally:
        an.close()


@pytest.mark.parametrize("system,name,tuning,half", [("vanderpol", "vanderpol_cells_acc4_backward.txt", None, 0.25),
                                                     ("springs", "springs_cells_acc10.txt.gz", "springs_reach_acc2.txt.gz", 0.125)])
def test_final_enclosures_restart_the_evolution(pkg, system, name, tuning, half):
    """ab200_reach_evolve_enclosures: evolving for T in one call and in two calls of T/2, the second started from the final
    TaylorSets of the first, gives final boxes that overlap the one-call ones with comparable widths -- and the models handed
    over decode to polynomials whose constant terms lie inside the final boxes. Springs (five variables) runs the build with
    two-word MultiIndex keys, whose records carry three words per term."""
    d = golden_io.load_dump(name)
    # (the cell boxes of the accuracy-10 springs golden with the tuning of the accuracy-2 one: no cell keys are formed here,
    # and accuracy 10 in five dimensions needs the domain-sized key layout of the analyser)
    ev = _evolver(pkg, golden_io.load_dump(tuning) if tuning else d, system=system, arena_words=262144, max_boxes=320)
    try:
        dim = ev.dims[0]
        cells = d["cells"][:8]
        locs = [ev.location(c["loc"]) for c in cells]
        boxes = pkg.pack_boxes([c["box"] for c in cells], dim)
        (_, _, _), (foff, fbox, _), enc = ev.reach_evolve_enclosures(locs, boxes, half, 5)
        assert len(enc) == int(foff[-1]) >= len(cells)
        assert np.array_equal(enc.unit_offsets, foff)
        for i in range(len(enc)):
            for k in range(dim):
                ex, co, err = enc.model(i, k)
                assert ex.shape[1] >= dim and err >= 0.0
                const = co[(ex.sum(axis=1) == 0)]
                c0 = const[0] if const.size else 0.0
                assert fbox[i, 2 * k] <= c0 <= fbox[i, 2 * k + 1]
        (_, _, _), (foff2, fbox2, _), enc2 = ev.reach_evolve_enclosures(None, None, half, 5, initial=enc)
        (_, _, _), (foff1, fbox1, _), _ = ev.reach_evolve_enclosures(locs, boxes, 2 * half, 5)
        assert len(enc2) == int(foff2[-1]) and foff2.shape[0] == len(enc) + 1
        # the same flow either way: where a cell gave one final set in every call, the two final boxes overlap (both enclose
        # the image of the cell under the flow for 0.5) and have comparable widths
        compared = 0
        for i in range(len(cells)):
            if foff[i + 1] - foff[i] != 1 or foff1[i + 1] - foff1[i] != 1:
                continue
            u = int(foff[i])
            if foff2[u + 1] - foff2[u] != 1:
                continue
            a = fbox1[int(foff1[i])]; b = fbox2[int(foff2[u])]
            compared += 1
            for k in range(dim):
                assert max(a[2 * k], b[2 * k]) <= min(a[2 * k + 1], b[2 * k + 1])
                w1 = a[2 * k + 1] - a[2 * k]; w2 = b[2 * k + 1] - b[2 * k]
                assert w2 <= 16 * w1 + 1e-9 and w1 <= 16 * w2 + 1e-9
        assert compared >= len(cells) // 2
        # the same TaylorSets rebuilt from their decoded models (ab200_enclosures_create) evolve to the same bits
        rebuilt = ev.enclosures_create([(int(enc.locs[i]), [enc.model(i, k) for k in range(dim)]) for i in range(len(enc))])
        (_, _, _), (foff3, fbox3, _), _ = ev.reach_evolve_enclosures(None, None, half, 5, initial=rebuilt)
        assert np.array_equal(foff3, foff2) and golden_io.boxes_equal_bits(fbox3, fbox2)
        with pytest.raises(pkg.Ab200Error):          # terms out of the Expansion order are refused
            ex0, co0, err0 = enc.model(0, 0)
            ev.enclosures_create([(int(enc.locs[0]), [(ex0[::-1], co0[::-1], err0)] * dim)])
        sub = enc.select([len(enc) - 1, 0])
        assert len(sub) == 2 and np.array_equal(sub.model(0, 0)[1], enc.model(len(enc) - 1, 0)[1])
    finally:
        ev.close()


@pytest.mark.parametrize("name,init", [("bouncingball_chain_acc5.txt", [0.0, 0.001, 1.999, 2.0]),
                                         ("bouncingball_chain_acc6.txt", [-2.0, 0.5, 0.5, 1.5]),
                                         ("bouncingball_backward_chain_acc5.txt", [2.0, 4.0, 0.0, 0.3])])
def test_outer_chain_reach_with_jumps_identical_cells(pkg, name, init):
    """BASELINE configs[1]: outer_chain_reach of the bouncing ball -- device laps (ignore_activations, urgent guard) plus
    the grid-side jump sets of src/reachability_restriction.cpp:340-375 between laps."""
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system("bouncingball"), [-10.1, 10.1, -0.1, 2.1], d["config"]["acc"], settings={"arena_words": 65536})
    try:
        assert an.lock_to_grid_time == d["config"]["lock"]
        reach = an.outer_chain_reach({"freefall": init}, backward="backward" in name)
        _, rb = an.enumerate(reach)
        assert golden_io.boxes_equal_bits(rb, d["sets"]["reach"]["freefall"])
        assert an.last_status.launches > 0
    finally:
        an.close()


HYBRID_CASES = [
    # (system, golden, domain, initial set, arena words)
    ("watertank", "watertank_reach_acc3.txt", [0.0, 1.0, 4.5, 9.0],
     {"flow,idle,rising": [1.0, 1.0, 7.6, 7.7], "flow,idle,falling": [0.0, 0.0, 5.7, 5.8]}, 131072),
    ("springs", "springs_reach_acc2.txt.gz", [0.0, 40.0, -2.0, 2.0, -3.0, 3.0, -1.0, 3.0, 0.0, 4.0],
     {"free1,free2": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0, 3.0]}, 131072),
    ("boost", "boost_reach_acc4.txt", [0.0, 0.001, 0.0, 3.0, 0.0, 6.0],
     {"incr,below_duty": [0.0, 0.0, 1.0, 1.0, 1.0, 1.0]}, 32768),
]


@pytest.mark.parametrize("system,name,domain,init,arena", HYBRID_CASES)
def test_upper_reach_evolve_hybrid_systems_identical_cells(pkg, system, name, domain, init, arena):
    """upper_reach_evolve with transitions taken inside the evolver (ignore_activations=false): blocking urgent jumps,
    activation windows of permissive / non-transverse events, sqrt and min() dynamics, several locations."""
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system(system), domain, d["config"]["acc"], settings={"arena_words": arena})
    try:
        assert an.lock_to_grid_time == d["config"]["lock"]
        reach, evolve = an.upper_reach_evolve(init, d["config"]["T"], d["config"]["events"])
        rl, rb = an.enumerate(reach)
        el, eb = an.enumerate(evolve)
        for which, locs, boxes in (("reach", rl, rb), ("evolve", el, eb)):
            for lname, ref in d["sets"][which].items():
                mine = boxes[locs == an.evolver.location(lname)]
                assert golden_io.boxes_equal_bits(mine, ref), "%s cells of %s in %s" % (which, system, lname)
    finally:
        an.close()


LOWER_CASES = [
    ("bouncingball", "bouncingball_lower_reach_acc4.txt", [-10.1, 10.1, -0.1, 2.1], {"freefall": [0.0, 0.5, 1.0, 2.0]}),
    ("vanderpol", "vanderpol_lower_reach_acc4.txt", [-3.0, 3.0, -3.0, 3.0], {"loc": [1.0, 1.5, 1.0, 1.6]}),
    ("watertank", "watertank_lower_reach_acc3.txt", [0.0, 1.0, 4.5, 9.0],
     {"flow,idle,rising": [1.0, 1.0, 7.6, 7.7], "flow,idle,falling": [0.0, 0.0, 5.7, 5.8]}),
]


@pytest.mark.parametrize("system,name,domain,init", LOWER_CASES)
def test_lower_reach_evolve_identical_cells(pkg, system, name, domain, init):
    """HybridReachabilityAnalyser::lower_reach_evolve (src/reachability_analyser.cpp:230-275)."""
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system(system), domain, d["config"]["acc"], settings={"arena_words": 131072, "max_boxes": 512})
    try:
        reach, evolve = an.lower_reach_evolve(init, d["config"]["T"], d["config"]["events"])
        rl, rb = an.enumerate(reach)
        el, eb = an.enumerate(evolve)
        for which, locs, boxes in (("reach", rl, rb), ("evolve", el, eb)):
            for lname, ref in d["sets"][which].items():
                mine = boxes[locs == an.evolver.location(lname)]
                assert golden_io.boxes_equal_bits(mine, ref), "%s cells of %s in %s" % (which, system, lname)
        assert an.evolver.settings.semantics == 0
    finally:
        an.close()


@pytest.mark.parametrize("name,init", [("bouncingball_chain_acc5.txt", [0.0, 0.001, 1.999, 2.0]),
                                       ("bouncingball_chain_acc6.txt", [-2.0, 0.5, 0.5, 1.5])])
def test_jump_set_checks_on_the_device_match_the_host_checker(pkg, name, init, monkeypatch):
    """forward_jump_set: the activation test of every cell box of the paving descent
    (ForwardDiscreteJumpSetChecker::check, src/set_checker.cpp:171-203) runs on the device, one warp per box per level
    (K_jump_check). AB200_JUMP_HOST=1 runs the same descent with the host build of the same check: both must give the
    golden's cells, and the device path must have launched its kernel."""
    d = golden_io.load_dump(name)
    results = []
    for host in (False, True):
        if host:
            monkeypatch.setenv("AB200_JUMP_HOST", "1")
        else:
            monkeypatch.delenv("AB200_JUMP_HOST", raising=False)
        an = pkg.HybridReachabilityAnalyser(_system("bouncingball"), [-10.1, 10.1, -0.1, 2.1], d["config"]["acc"], settings={"arena_words": 65536})
        try:
            reach = an.outer_chain_reach({"freefall": init})
            _, rb = an.enumerate(reach)
            assert golden_io.boxes_equal_bits(rb, d["sets"]["reach"]["freefall"])
            results.append(reach.keys.copy())
        finally:
            an.close()
    assert np.array_equal(results[0], results[1])


@pytest.mark.parametrize("system,name,wpc", [("bouncingball", "bouncingball_cells_acc5.txt", 8), ("bouncingball", "bouncingball_cells_acc5.txt", 3),
                                              ("watertank", "watertank_cells_acc6.txt.gz", 6), ("vanderpol", "vanderpol_cells_acc7.txt.gz", 12)])
def test_phase_aligned_ctas_with_events_bit_exact(pkg, system, name, wpc, monkeypatch):
    """CTAs of several warps that meet at the stages of a step (the launch shape of large frontiers), forced onto small hybrid
    batches: warps of one CTA then take different paths (urgent bounces, permissive jumps with activation windows, early
    termination, cells of different length, warps running out of cells), which is exactly what the alignment protocol must
    survive without waiting for ever. Results are the goldens, bit for bit."""
    monkeypatch.setenv("AB200_WPC", str(wpc))
    d = golden_io.load_dump(name)
    ev = _evolver(pkg, d, system=system, arena_words=131072, max_boxes=2048)
    try:
        dim = ev.dims[0]
        locs = [ev.location(c["loc"]) for c in d["cells"]]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
        events = 1 if d["config"]["chain"] else d["config"]["events"]
        (roff, rbox, rloc), (foff, fbox, floc) = ev.reach_evolve(locs, boxes, d["config"]["T"], events, d["config"]["ignore_activations"])
        for i, c in enumerate(d["cells"]):
            for mine_b, mine_l, ref_b, ref_l in ((rbox[roff[i]:roff[i + 1]], rloc[roff[i]:roff[i + 1]], c["reach"], c["reach_loc"]),
                                                 (fbox[foff[i]:foff[i + 1]], floc[foff[i]:foff[i + 1]], c["final"], c["final_loc"])):
                assert len(ref_b) == mine_b.shape[0]
                for lname in set(ref_l):
                    ref_rows = [b for b, n in zip(ref_b, ref_l) if n == lname]
                    assert golden_io.boxes_equal_bits(mine_b[mine_l == ev.location(lname), :2 * dim], ref_rows), "cell %d in %s" % (c["id"], lname)
    finally:
        ev.close()


def test_boost_outer_chain_reach_fails_like_the_reference(pkg):
    """configs[4]: the reference's own outer_chain_reach on the boost converter throws in forward_jump_set because
    lie_derivative cannot differentiate the min() guard (src/function.cpp:1329). Same laps, same failure here."""
    an = pkg.HybridReachabilityAnalyser(_system("boost"), [0.0, 0.001, 0.0, 3.0, 0.0, 6.0], 1, settings={"arena_words": 524288})
    try:
        with pytest.raises(pkg.Ab200Error) as ei:
            an.outer_chain_reach({"incr,below_duty": [0.0, 0.0, 1.0, 1.0, 1.0, 1.0]})
        assert "Lie derivative" in str(ei.value)
        assert an.last_status.launches > 0
    finally:
        an.close()


def test_lap_cell_sets_match_union_of_reference_discretisations(pkg):
    d = golden_io.load_dump("vanderpol_cells_acc3.txt")
    ev = _evolver(pkg, d, arena_words=131072)
    try:
        dim = ev.dims[0]
        locs = [ev.location(c["loc"]) for c in d["cells"]]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
        reach, evolve = ev.upper_reach_evolve(locs, boxes, d["config"]["T"], d["config"]["events"])
        # sortedness / uniqueness of the emitted keys
        assert np.all(reach.keys[1:] > reach.keys[:-1]) and np.all(evolve.keys[1:] > evolve.keys[:-1])
        # leaf-cell counts equal the area of the reference's merged cells
        g = d["grid"]["loc"]; acc = d["config"]["acc"]
        def leaves(cells):
            tot = 0
            for b in cells:
                tot += round((b[1] - b[0]) / (g["lengths"][0] / 2 ** acc)) * round((b[3] - b[2]) / (g["lengths"][1] / 2 ** acc))
            return tot
        assert len(reach) == leaves(d["sets"]["ureach"]["loc"])
        assert len(evolve) == leaves(d["sets"]["ufinal"]["loc"])
    finally:
        ev.close()


DEPTH_CASES = [
    # BASELINE.json grid depths: configs[0] watertank depth 6, configs[3] springs depth 10, configs[4] boost at accuracy 6
    ("watertank", "watertank_cells_acc6.txt.gz", [0.0, 1.0, 4.5, 9.0], 131072),
    ("springs", "springs_cells_acc10.txt.gz", [0.0, 40.0, -2.0, 2.0, -3.0, 3.0, -1.0, 3.0, 0.0, 4.0], 262144),
    ("boost", "boost_cells_acc6.txt.gz", [0.0, 0.001, 0.0, 3.0, 0.0, 6.0], 32768),
]


@pytest.mark.parametrize("system,name,domain,arena", DEPTH_CASES)
def test_baseline_depth_sampled_cells_bit_exact(pkg, system, name, domain, arena):
    """The three hybrid configs at the grid depth BASELINE.json quotes: sampled frontier cells over a horizon the compiled
    reference finishes in seconds. Per-enclosure reach / final boxes bit-exact, and the discretised union (what
    UpperReachEvolveWorker adjoins) identical cell for cell (count + sha256 of the sorted cell boxes per location).
    springs at accuracy 10 needs the domain-sized key layout (ab200_set_accuracy_for_domain)."""
    import sys as _sys
    _sys.path.insert(0, golden_io.GOLDEN)
    from bench_fingerprint import fingerprint
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system(system), domain, d["config"]["acc"], settings={"arena_words": arena, "max_boxes": 2048, "hash_log2": 22})
    ev = an.evolver
    try:
        assert an.lock_to_grid_time == d["config"]["lock"]
        for lname, g in d["grid"].items():
            l = ev.location(lname)
            assert list(an.grid_lengths[l][:g["n"]]) == g["lengths"] and list(an.grid_origin[l][:g["n"]]) == g["origin"]
            ev.tune_settings(l, g["maxstep"], g["refw"], g["origin"], g["lengths"])
        dim = ev.dims[0]
        locs = [ev.location(c["loc"]) for c in d["cells"]]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
        T, events = d["config"]["T"], d["config"]["events"]
        (roff, rbox, rloc), (foff, fbox, floc) = ev.reach_evolve(locs, boxes, T, events, d["config"]["ignore_activations"])
        for i, c in enumerate(d["cells"]):
            assert c["error"] is None
            # the reference's list sets are grouped by location (HybridListSet is a map over locations, iterated in name
            # order); within a location the insertion order is kept. Compare location by location.
            for mine_b, mine_l, ref_b, ref_l, what in ((rbox[roff[i]:roff[i + 1]], rloc[roff[i]:roff[i + 1]], c["reach"], c["reach_loc"], "reach"),
                                                       (fbox[foff[i]:foff[i + 1]], floc[foff[i]:foff[i + 1]], c["final"], c["final_loc"], "final")):
                assert len(ref_b) == mine_b.shape[0], "%s box count of cell %d" % (what, c["id"])
                for lname in set(ref_l) | set(d["grid"].keys()):
                    ref_rows = [b for b, n in zip(ref_b, ref_l) if n == lname]
                    assert golden_io.boxes_equal_bits(mine_b[mine_l == ev.location(lname), :2 * dim], ref_rows), "%s boxes of cell %d in %s" % (what, c["id"], lname)
        # sqrt / rec series: how many calls could have changed with the last bits of libm's log / pow (0 = certified exact)
        print("%s: %d steps, %d libm-sensitive series calls" % (system, ev.last_status.steps, ev.last_status.libm_sensitive))
        assert ev.last_status.libm_sensitive <= ev.last_status.steps // 100
        if system != "watertank":
            assert ev.last_status.libm_sensitive == 0          # no sqrt / rec of a non-constant model in these systems
        reach, evolve = ev.upper_reach_evolve(locs, boxes, T, events, d["config"]["ignore_activations"])
        for which, cs in (("ureach", reach), ("ufinal", evolve)):
            cl, cb = an.enumerate(cs)
            for lname, want in d["fp"][which].items():
                mine = fingerprint(cb[cl == ev.location(lname)])
                assert mine["cells"] == want["cells"] and (want["cells"] == 0 or mine["sha256"] == want["sha256"]), "%s cells of %s in %s" % (which, system, lname)
    finally:
        an.close()


CONSTRAINT_CASES = [
    ("vanderpol", "vanderpol_cons_disc_acc4.txt", [-3.0, 3.0, -3.0, 3.0], 131072),
    ("vanderpol", "vanderpol_cons_band_acc6.txt", [-3.0, 3.0, -3.0, 3.0], 32768),
    ("bouncingball", "bouncingball_cons_disc_chain_acc5.txt", [-10.1, 10.1, -0.1, 2.1], 65536),
]


@pytest.mark.parametrize("system,name,domain,arena", CONSTRAINT_CASES)
def test_constraint_set_initial_sets_identical_cells(pkg, system, name, domain, arena):
    """Initial sets given as constraint sets { x in box : f(x) in codomain } (HybridBoundedConstraintSet): the initial cells
    (outer_intersection_with: ConstraintSet::overlaps by subdivision, covers by one evaluation, decided cell by cell down the
    paving) and the upper_reach_evolve / outer_chain_reach run from them, against the compiled reference."""
    d = golden_io.load_dump(name)
    an = pkg.HybridReachabilityAnalyser(_system(system), domain, d["config"]["acc"], settings={"arena_words": arena})
    try:
        initial = an.discretise_constraint("".join(d["cons_text"]))
        il, ib = an.enumerate(initial)
        for lname, ref in d["sets"]["initial"].items():
            assert golden_io.boxes_equal_bits(ib[il == an.evolver.location(lname)], ref), "initial cells in %s" % lname
        assert len(ib) == sum(len(v) for v in d["sets"]["initial"].values()) > 0
        if d["config"]["chain"]:
            reach = an.outer_chain_reach_from(initial)
            sets = {"reach": reach}
        else:
            reach, evolve = an.upper_reach_evolve_from(initial, d["config"]["T"], d["config"]["events"])
            sets = {"reach": reach, "evolve": evolve}
        for which, cs in sets.items():
            cl, cb = an.enumerate(cs)
            for lname, ref in d["sets"][which].items():
                assert golden_io.boxes_equal_bits(cb[cl == an.evolver.location(lname)], ref), "%s cells in %s" % (which, lname)
    finally:
        an.close()


def test_bench_lap_matches_the_reference_fingerprint(pkg):
    """The benchmark's own workload at full size: van der Pol, accuracy 7, all 101,520 frontier cells, one lock-to-grid lap
    (the lane build). Its reach / evolve sets must be, leaf cell for leaf cell, what the compiled reference's
    UpperReachEvolveWorker::get_result() returns on the same frontier (tests/golden/vanderpol_bench_fingerprint.json, 13.5
    CPU-hours of the reference in six chunks, tests/golden/make_bench_fingerprint.sh)."""
    import json
    import sys as _sys
    _sys.path.insert(0, golden_io.GOLDEN)
    from bench_fingerprint import leaf_fingerprint
    want = json.load(open(os.path.join(golden_io.GOLDEN, "vanderpol_bench_fingerprint.json")))
    an = pkg.HybridReachabilityAnalyser(_system(), [-3.0, 3.0, -3.0, 3.0], 7, settings={"arena_words": 32768, "max_boxes": 320, "hash_log2": 25})
    try:
        initial = an.discretise({0: [-2.2, 2.2, -2.8, 2.8]})
        _, cells = an.enumerate(initial, mince=True)
        assert cells.shape[0] == want["frontier_cells"] == 101520
        an.upper_reach_evolve({0: [1.0, 1.0 + 1e-3, 1.0, 1.0 + 1e-3]}, 1e-3, 5)            # the tuning upload of a lap
        ev = an.evolver
        reach, evolve = ev.upper_reach_evolve(np.zeros((cells.shape[0],), dtype=np.int32), pkg.pack_boxes(cells, 2), an.lock_to_grid_time, 5)
        assert ev.last_status.build == 2
        for name, cs in (("reach", reach), ("evolve", evolve)):
            got = leaf_fingerprint(ev.decode(cs.keys)[:, 1:3])
            assert got["cells"] == want[name]["cells"] and got["sha256"] == want[name]["sha256"], name
    finally:
        an.close()


def test_flow_bounds_facade(pkg):
    d = golden_io.load_dump("vanderpol_cells_acc6.txt.gz")
    ev = _evolver(pkg, d)
    try:
        dim = ev.dims[0]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], dim)
        h, b, s = ev.flow_bounds([0] * len(d["cells"]), boxes, d["grid"]["loc"]["maxstep"])
        assert np.all(s == 0) and np.all(h > 0) and np.all(h <= d["grid"]["loc"]["maxstep"])
        # the bounding box contains the starting box (ARIADNE_ASSERT(subset(d,b)), src/taylor_calculus.cpp:125)
        assert np.all(b[:, 0] <= boxes[:, 0]) and np.all(b[:, 1] >= boxes[:, 1])
    finally:
        ev.close()


def test_edge_cases_and_error_codes(pkg):
    """Empty frontier, first-failing-cell error reporting (include/workers.h:129-139,159-166) and device capacity limits."""
    d = golden_io.load_dump("vanderpol_cells_acc3.txt")
    ev = _evolver(pkg, d, arena_words=131072)
    try:
        T, events = d["config"]["T"], d["config"]["events"]
        # empty frontier -> two empty sets, no launch of the evolution kernel needed
        r, e = ev.upper_reach_evolve(np.zeros((0,), dtype=np.int32), np.zeros((0, pkg.BOX_STRIDE)), T, events)
        assert len(r) == 0 and len(e) == 0
        cells = d["cells"][:3]
        locs = np.zeros((len(cells),), dtype=np.int32)
        boxes = pkg.pack_boxes([c["box"] for c in cells], 2)
        # an unknown location is rejected before anything is launched
        with pytest.raises(pkg.Ab200Error) as ei:
            ev.upper_reach_evolve(np.array([0, 7, 0], dtype=np.int32), boxes, T, events)
        assert ei.value.code == pkg.AB200_E_ARGUMENT
        # a scratch arena that is too small fails the lap with the index of the first failing cell, no partial result
        ev.set_settings(arena_words=512)
        with pytest.raises(pkg.Ab200Error) as ei:
            ev.upper_reach_evolve(locs, boxes, T, events)
        assert ei.value.code == pkg.AB200_E_ARENA and ev.last_status.cell == 0
        # too few output boxes per cell
        ev.set_settings(arena_words=131072, max_boxes=2)
        with pytest.raises(pkg.Ab200Error) as ei:
            ev.upper_reach_evolve(locs, boxes, T, events)
        assert ei.value.code == pkg.AB200_E_OUTPUT
        # and the same context still works afterwards
        ev.set_settings(max_boxes=192)
        r, e = ev.upper_reach_evolve(locs, boxes, T, events)
        assert len(r) > 0 and len(e) > 0
        # a degenerate (point) cell is a valid enclosure (the bounder widens singleton coordinates by 1e-16)
        r1, e1 = ev.upper_reach_evolve(np.zeros((1,), dtype=np.int32), pkg.pack_boxes([[1.0, 1.0, 1.0, 1.0]], 2), T, events)
        assert len(r1) > 0
    finally:
        ev.close()


def test_round_trip_properties_at_scale(pkg):
    """Size-independent properties on a frontier the oracle could not finish in test time: (1) the lap of a union of
    frontiers is the union of the laps (cells are independent units); (2) evolving a frontier in one call or in small
    batches gives identical key arrays; (3) keys come back sorted and duplicate-free; (4) every frontier cell is
    contained in the reach set of its own lap (the flow tube starts on the cell)."""
    an = pkg.HybridReachabilityAnalyser(_system(), [-3.0, 3.0, -3.0, 3.0], 6, settings={"arena_words": 16384})
    try:
        own = an.discretise({"loc": [-0.8, 0.8, -0.8, 0.8]})
        _, cells = an.enumerate(own, mince=True)
        n = cells.shape[0]
        assert n == len(own) and n > 2000
        locs = np.zeros((n,), dtype=np.int32)
        boxes = pkg.pack_boxes(cells, 2)
        an.upper_reach_evolve({"loc": [1.0, 1.001, 1.0, 1.001]}, 1e-3, 5)       # uploads the evolver tuning of a lap
        T = an.lock_to_grid_time / 4
        ev = an.evolver
        r_all, e_all = ev.upper_reach_evolve(locs, boxes, T, 5)
        assert np.all(np.diff(r_all.keys) > 0) and np.all(np.diff(e_all.keys) > 0)
        half = n // 2
        r_a, e_a = ev.upper_reach_evolve(locs[:half], boxes[:half], T, 5)
        r_b, e_b = ev.upper_reach_evolve(locs[half:], boxes[half:], T, 5)
        assert np.array_equal(np.union1d(r_a.keys, r_b.keys), r_all.keys)
        assert np.array_equal(np.union1d(e_a.keys, e_b.keys), e_all.keys)
        ev.set_settings(batch_cells=256)
        r_c, e_c = ev.upper_reach_evolve(locs, boxes, T, 5)
        assert np.array_equal(r_c.keys, r_all.keys) and np.array_equal(e_c.keys, e_all.keys)
        assert np.isin(own.keys, r_all.keys).all()
    finally:
        an.close()


def test_lane_build_matches_warp_build(pkg, monkeypatch):
    """The one-lane-per-cell build (used above AB_LANE_THRESHOLD cells) and the one-warp-per-cell build are two schedules of
    the same arithmetic: identical boxes and identical cell keys on the same frontier. Also checks the golden boxes through
    the lane build, and that an unknown AB200_TUNE value is refused."""
    d = golden_io.load_dump("vanderpol_cells_acc7.txt.gz")
    ev = _evolver(pkg, d, arena_words=32768, max_boxes=320)
    try:
        locs = [ev.location(c["loc"]) for c in d["cells"]]
        boxes = pkg.pack_boxes([c["box"] for c in d["cells"]], 2)
        T, events = d["config"]["T"], d["config"]["events"]
        out = {}
        for build in ("warp", "lane"):
            monkeypatch.setenv("AB200_TUNE", build)
            (roff, rbox, rloc), (foff, fbox, floc) = ev.reach_evolve(locs, boxes, T, events, d["config"]["ignore_activations"])
            assert ev.last_status.build == (2 if build == "lane" else 0)
            r, e = ev.upper_reach_evolve(np.array(locs, dtype=np.int32), boxes, T, events)
            out[build] = (roff.copy(), rbox.copy(), foff.copy(), fbox.copy(), r.keys.copy(), e.keys.copy())
            for i, c in enumerate(d["cells"]):
                assert golden_io.boxes_equal_bits(rbox[roff[i]:roff[i + 1], :4], c["reach"]), "%s build, reach boxes of cell %d" % (build, c["id"])
                assert golden_io.boxes_equal_bits(fbox[foff[i]:foff[i + 1], :4], c["final"]), "%s build, final boxes of cell %d" % (build, c["id"])
        for a, b in zip(out["warp"], out["lane"]):
            assert a.shape == b.shape and a.tobytes() == b.tobytes()
        monkeypatch.setenv("AB200_TUNE", "no-such-build")
        with pytest.raises(pkg.Ab200Error):
            ev.upper_reach_evolve(np.array(locs, dtype=np.int32), boxes, T, events)
    finally:
        ev.close()


ADAPTER_CASES = [
    ("vanderpol", "vanderpol_reach_acc4.txt", ["acc=4", "init=-1.6,-1.3,0.3,0.8", "laps=1"]),
    ("vanderpol", "vanderpol_reach_acc3.txt", ["acc=3", "init=1.0,1.3,1.0,1.6", "laps=2"]),
    ("bouncingball", "bouncingball_chain_acc5.txt", ["acc=5"]),
    ("watertank", "watertank_reach_acc3.txt", ["acc=3", "T=0.5", "initr=1.0,1.0,7.6,7.7", "initf=0.0,0.0,5.7,5.8"]),
    ("boost", "boost_reach_acc4.txt", ["acc=4", "chain=0", "T=0.011"]),
]


@pytest.mark.parametrize("config,golden,args", ADAPTER_CASES)
def test_reference_analyser_with_b200_worker(tmp_path, config, golden, args):
    """The drop-in boundary, compiled: the REFERENCE's own HybridReachabilityAnalyser (unmodified sources, oracle/_ref) with
    adapter/workers_b200 swapped in where src/reachability_analyser.cpp:192-198 builds its UpperReachEvolveWorker. Every lap
    goes through ab200_upper_reach_evolve; everything else (grid, lock time, restriction, jump sets, BDD set algebra) is the
    reference's. Its result must be the stock run's (the committed golden) cell for cell."""
    import subprocess
    exe = os.path.join(os.path.dirname(golden_io.GOLDEN), "..", "oracle", "_ref", "ref_harness_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_harness_b200 has not been built (make -C oracle adapter)")
    out = str(tmp_path / "b200.txt")
    subprocess.check_call([exe, config, "reach", "out=" + out] + args, timeout=600)
    mine = golden_io.load_dump(out)
    assert "reach" in mine["sets"], open(out).read()[-600:]
    ref = golden_io.load_dump(golden)
    assert mine["config"]["lock"] == ref["config"]["lock"]
    for which in ref["sets"]:                 # (a chain-reach dump holds the reach set only)
        assert set(mine["sets"][which].keys()) == set(ref["sets"][which].keys())
        for lname, cells in ref["sets"][which].items():
            assert golden_io.boxes_equal_bits(np.array(mine["sets"][which][lname]), cells), "%s cells of %s in %s" % (which, config, lname)


ADAPTER_EPS_CASES = [
    ("bouncingball", "bouncingball_epslower_acc4.txt", ["acc=4"]),
    ("vanderpol", "vanderpol_epslower_acc3.txt", ["acc=3", "prune=0", "init=1.0,1.3,1.0,1.6"]),
    ("watertank", "watertank_epslower_acc3_prune.txt", ["acc=3", "prune=1"]),
]


@pytest.mark.parametrize("config,golden,args", ADAPTER_EPS_CASES)
def test_reference_analyser_with_b200_lower_worker(tmp_path, config, golden, args):
    """The reference's own epsilon_lower_chain_reach (unmodified src/reachability_analyser.cpp:497-627: parameter midpoints,
    restriction, pruning with its own rand(), epsilon bookkeeping) with adapter/workers_b200's B200LowerReachEpsilonWorker
    where :583 builds its LowerReachEpsilonWorker: the TaylorSets of every lap cross the C ABI both ways
    (ab200_enclosures_create / ab200_enclosures_model). Cells and epsilon must equal the stock run's bit for bit."""
    import subprocess
    exe = os.path.join(os.path.dirname(golden_io.GOLDEN), "..", "oracle", "_ref", "ref_harness_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_harness_b200 has not been built (make -C oracle adapter)")
    out = str(tmp_path / "b200_eps.txt")
    subprocess.check_call([exe, config, "epslower", "out=" + out] + args, timeout=600)
    mine = golden_io.load_dump(out); ref = golden_io.load_dump(golden)
    assert "reach" in mine["sets"], open(out).read()[-600:]
    assert set(mine["sets"]["reach"].keys()) == set(ref["sets"]["reach"].keys())
    for lname, cells in ref["sets"]["reach"].items():
        assert golden_io.boxes_equal_bits(np.array(mine["sets"]["reach"][lname]), cells), "reach cells of %s in %s" % (config, lname)
    assert set(mine["epsilon"].keys()) == set(ref["epsilon"].keys())
    for lname, e in ref["epsilon"].items():
        assert golden_io.boxes_equal_bits(mine["epsilon"][lname], e), "epsilon of %s in %s" % (config, lname)


EVOLVER_CASES = [
    ("vanderpol", "vanderpol_cells_acc6.txt.gz", ["acc=6", "mince=1", "ncells=24", "stride=1050"]),
    ("bouncingball", "bouncingball_cells_acc5.txt", ["acc=5", "init=-3.0,3.0,0.0,2.0", "ncells=60", "stride=37"]),
    ("watertank", "watertank_cells_acc6.txt.gz", ["acc=6", "T=3.0", "ncells=24", "stride=14"]),
]


@pytest.mark.parametrize("config,golden,args", EVOLVER_CASES)
def test_reference_evolver_interface_with_b200_evolution(tmp_path, config, golden, args):
    """The EvolverInterface seam, compiled (adapter/evolver_b200.h): an ImageSetHybridEvolver whose _evolution runs on the
    device, driven by the reference's own reach_evolve(enclosure, time, ignore_activations, direction, semantics). Every
    enclosure it returns must contain the stock evolver's (north_star: "every per-cell final enclosure must contain the
    reference enclosure"), with the same number of enclosures per location, and be no wider than a few ulps more."""
    import subprocess
    exe = os.path.join(os.path.dirname(golden_io.GOLDEN), "..", "oracle", "_ref", "ref_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_harness has not been built (make -C oracle harness)")
    out = str(tmp_path / "b200_evolver.txt")
    subprocess.check_call([exe, config, "cells", "b200=1", "out=" + out] + args, timeout=600)
    mine = golden_io.load_dump(out); ref = golden_io.load_dump(golden)
    assert len(mine["cells"]) == len(ref["cells"]) > 0
    n_boxes = 0
    for cm, cr in zip(mine["cells"], ref["cells"]):
        assert cm["error"] is None and cm["id"] == cr["id"]
        for which in ("reach", "final"):
            assert len(cm[which]) == len(cr[which]), "%s enclosures of cell %d" % (which, cr["id"])
            for lname in set(cr[which + "_loc"]):
                a = np.array([b for b, n in zip(cm[which], cm[which + "_loc"]) if n == lname])
                b = np.array([b for b, n in zip(cr[which], cr[which + "_loc"]) if n == lname])
                assert a.shape == b.shape, "%s enclosures of cell %d in %s" % (which, cr["id"], lname)
                lo_m, hi_m, lo_r, hi_r = a[:, 0::2], a[:, 1::2], b[:, 0::2], b[:, 1::2]
                assert np.all(lo_m <= lo_r) and np.all(hi_m >= hi_r), "containment, cell %d" % cr["id"]
                tol = 1e-12 * np.maximum(hi_r - lo_r, 0.0) + 64 * np.spacing(np.maximum(np.abs(lo_r), np.abs(hi_r)))
                assert np.all(lo_r - lo_m <= tol) and np.all(hi_m - hi_r <= tol), "excess width, cell %d" % cr["id"]
                n_boxes += a.shape[0]
    assert n_boxes > 0


def _load_integrator_golden(name):
    out = {"models": []}
    with open(os.path.join(golden_io.GOLDEN, name)) as f:
        for line in f:
            t = line.split()
            if t[0] == "integrator":
                out["loc"] = t[2]; out["order"] = int(t[4]); out["hmax"] = float.fromhex(t[6]); out["box"] = [float.fromhex(x) for x in t[8:]]
            elif t[0] == "bounds":
                out["h"] = float.fromhex(t[2]); out["bounds"] = [float.fromhex(x) for x in t[4:]]
            elif t[0] == "domain":
                out["domain"] = [float.fromhex(x) for x in t[1:]]
            elif t[0] == "model":
                out["models"].append({"error": float.fromhex(t[5]), "exps": [], "coefs": []})
            elif t[0] == "t":
                out["models"][-1]["exps"].append([int(x) for x in t[1:-1]]); out["models"][-1]["coefs"].append(float.fromhex(t[-1]))
    return out


@pytest.mark.parametrize("system,name", [("constder", "integrator_constder_o1.txt"), ("constder", "integrator_constder.txt"),
                                         ("vanderpol", "integrator_vanderpol_o4.txt"), ("vanderpol", "integrator_vanderpol_o6.txt"),
                                         ("bouncingball", "integrator_bouncingball.txt")])
def test_integrator_flow_bounds_and_flow_bit_exact(pkg, system, name):
    """IntegratorInterface: IntegratorBase::flow_bounds with the integrator's own constants and
    TaylorIntegrator::flow (src/integrator.cpp:44-152), against the reference's results on the vector field, box and step of
    test/test_integrator.cpp:73-157 (x'=2, y'=3, box [0,1]x[-0.5,1.5], h=1, order 1) and on the van der Pol / bouncing-ball
    dynamics: step, bounding box, and every term, coefficient and error of the flow's Taylor models, bit for bit."""
    g = _load_integrator_golden(name)
    ev = pkg.Evolver(_system(system), settings={"arena_words": 262144})
    try:
        loc = ev.location(g["loc"])
        h, b = ev.integrator_flow_bounds(loc, g["box"], g["hmax"])
        assert h == g["h"] and golden_io.boxes_equal_bits(b, g["bounds"])
        h2, dom, models = ev.integrator_flow(loc, g["box"], g["hmax"], g["order"])
        assert h2 == g["h"] and golden_io.boxes_equal_bits(dom, g["domain"])
        assert len(models) == len(g["models"])
        for (ex, cf, err), ref in zip(models, g["models"]):
            assert ex.tolist() == ref["exps"]
            assert golden_io.boxes_equal_bits(cf, ref["coefs"]) and golden_io.boxes_equal_bits([err], [ref["error"]])
        if system == "constder":
            # the reference test's own checks (test/test_integrator.cpp:73-107): step kept, bounds cover [0,3]x[-0.5,4.5], and the
            # flow is (x0+2t, y0+3t) to 1e-10 in the supremum norm of the scaled polynomial
            assert h == 1.0 and b[0] <= 0.0 and b[1] >= 3.0 and b[2] <= -0.5 and b[3] >= 4.5
            expect = [{(0, 0, 0): 0.5, (1, 0, 0): 0.5, (0, 0, 1): 2.0}, {(0, 0, 0): 0.5, (0, 1, 0): 1.0, (0, 0, 1): 3.0}]
            for (ex, cf, err), want in zip(models, expect):
                got = {tuple(e): c for e, c in zip(ex.tolist(), cf.tolist())}
                norm = err + sum(abs(got.get(k, 0.0) - want.get(k, 0.0)) for k in set(got) | set(want))
                assert norm < 1e-10
    finally:
        ev.close()
