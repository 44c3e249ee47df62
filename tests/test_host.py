"""CPU-side checks (run with -m "not gpu"): golden vectors are well formed, the C-ABI library loads and exports every
symbol that include/ariadne_b200.h declares, and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

import golden_io


def test_header_symbols_exported(pkg, root):
    header = open(os.path.join(root, "include", "ariadne_b200.h")).read()
    declared = set(re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", header))
    declared -= {"ab200_ctx", "ab200_cellset", "ab200_boxlist", "ab200_analyser", "ab200_settings", "ab200_status", "ab200_enclosures"}
    assert declared == set(pkg.EXPORTS), (declared ^ set(pkg.EXPORTS))
    lib = ctypes.CDLL(pkg.library_path())
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_default_settings_match_reference(pkg):
    lib = pkg.load_library()
    s = pkg.Settings()
    lib.ab200_default_settings(ctypes.byref(s))
    # TaylorCalculus defaults src/taylor_calculus.cpp:136-146; evolver defaults src/hybrid_evolver-image.cpp:1347-1359
    assert (s.temporal_order, s.spacial_order) == (6, 6)
    assert s.sweep_threshold == 1e-12 and s.maximum_enclosure_widths_ratio == 5.0


def test_no_cpu_fallback(pkg):
    """Without a GPU the context cannot be created and nothing computes."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.Ab200Error) as e:
        pkg.Evolver(open(os.path.join(golden_io.GOLDEN, "vanderpol.sys")).read())
    assert e.value.code == 100


def test_golden_vectors_well_formed():
    d = golden_io.load_dump("vanderpol_cells_acc6.txt.gz")
    assert len(d["cells"]) == 24 and d["config"]["acc"] == 6
    for c in d["cells"]:
        assert c["error"] is None and len(c["reach"]) >= 1 and len(c["final"]) == 1
        assert len(c["dreach"]) >= 1
    r = golden_io.load_dump("vanderpol_reach_acc3.txt")
    assert len(r["sets"]["reach"]["loc"]) == 22 and len(r["sets"]["evolve"]["loc"]) == 15


def test_system_texts_parse_on_the_host(pkg):
    """The op-tape parser (release-1.0_b200/csrc/ab200_system.cpp) runs without a device: every system of the five
    BASELINE configs is accepted with the expected shape, malformed input is rejected with a message."""
    expect = {"vanderpol": (1, 0), "bouncingball": (1, 1), "watertank": (4, 8), "springs": (2, 2), "boost": (3, 4)}
    for name, (nloc, nev) in expect.items():
        text = open(os.path.join(golden_io.GOLDEN, name + ".sys")).read()
        got = pkg.system_check(text)
        assert got[:2] == (nloc, nev) and got[2] > 0, (name, got)
    with pytest.raises(pkg.Ab200Error):
        pkg.system_check("system broken\nlocations 1\nlocation 0 a n 2 events 0\n  dyn 0 V7 ;\n")
    with pytest.raises(pkg.Ab200Error):
        pkg.system_check("system broken\nlocations 1\nlocation 0 a n 1 events 0\n  dyn 0 V0 FROB ;\n")
    # events out of name order: index order would no longer be the reference's std::set<DiscreteEvent> order
    wt = open(os.path.join(golden_io.GOLDEN, "watertank.sys")).read()
    lines = wt.split("\n")
    ev = [i for i, l in enumerate(lines) if l.strip().startswith("event ")]
    names = [lines[i].split()[2] for i in ev[:2]]
    swapped = list(lines)
    swapped[ev[0]] = lines[ev[0]].replace(" " + names[0] + " ", " " + names[1] + " ", 1)
    swapped[ev[1]] = lines[ev[1]].replace(" " + names[1] + " ", " " + names[0] + " ", 1)
    with pytest.raises(pkg.Ab200Error) as e:
        pkg.system_check("\n".join(swapped))
    assert "ascending name order" in str(e.value)
