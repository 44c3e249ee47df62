"""The oracle is the compiled, unmodified reference (oracle/_ref/ref_harness, built by `make -C oracle all harness` from
the sources under /root/reference; the reference's own unit tests run against the same library with `make -C oracle
reftests`). These CPU tests pin it to the committed golden vectors: regenerating a golden must give the same bytes, so
the fixtures the GPU parity tests compare against are exactly what the reference computes."""
import os
import subprocess

import pytest

import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

pytestmark = pytest.mark.skipif(not os.path.exists(HARNESS), reason="compiled reference (oracle/_ref) not built")


def _lines(text):
    return [l for l in text.splitlines() if not l.startswith("seconds ")]


def _run(*args):
    return subprocess.run([HARNESS] + list(args), stdout=subprocess.PIPE, timeout=600, check=True).stdout.decode()


@pytest.mark.parametrize("name", ["vanderpol", "bouncingball", "watertank", "springs", "boost", "pendulum"])
def test_system_tapes_are_what_the_reference_automata_hold(name):
    assert _run(name, "system") == open(os.path.join(golden_io.GOLDEN, name + ".sys")).read()


@pytest.mark.parametrize("golden,args", [
    ("bouncingball_chain_acc5.txt", ["bouncingball", "reach", "acc=5"]),
    ("vanderpol_reach_acc4.txt", ["vanderpol", "reach", "acc=4", "init=-1.6,-1.3,0.3,0.8", "laps=1"]),
    ("bouncingball_cells_acc4.txt", ["bouncingball", "cells", "acc=4", "init=0.0,0.5,1.0,2.0"]),
    ("boost_reach_acc4.txt", ["boost", "reach", "acc=4", "chain=0", "T=0.011"]),
])
def test_oracle_reproduces_committed_goldens(golden, args):
    assert _lines(_run(*args)) == _lines(golden_io._open(golden).read())


@pytest.mark.parametrize("golden,args", [
    ("bouncingball_epslower_acc4.txt", ["bouncingball", "epslower", "acc=4"]),
    ("springs_epslower_acc2_lock1.txt", ["springs", "epslower", "acc=2", "prune=1", "lock=1.0"]),
    ("watertank_epslower_acc3_prune.txt", ["watertank", "epslower", "acc=3", "prune=1"]),
])
def test_oracle_reproduces_epsilon_lower_chain_reach_goldens(golden, args, monkeypatch):
    """epsilon_lower_chain_reach with pruning draws rand() in the order its worker threads finish: one thread (ORACLE_THREADS=1,
    oracle/shim/boost/thread.hpp) gives the reproducible run the goldens were taken from."""
    monkeypatch.setenv("ORACLE_THREADS", "1")
    assert _lines(_run(*args)) == _lines(golden_io._open(golden).read())


def test_midpoint_parameter_system_is_what_the_reference_substitutes():
    assert _run("watertank", "system", "midparams=1") == open(os.path.join(golden_io.GOLDEN, "watertank_midparams.sys")).read()


def test_primitive_kat_vectors_come_from_the_reference_and_hold_its_unit_test_answers():
    """tests/golden/kat_primitives.txt.gz is the output of oracle/_ref/ref_kat (the reference library itself); it must
    regenerate byte for byte, and it must contain the literal answers the reference's own unit tests assert
    (test/test_taylor_model.cpp:176,231-234; test/test_interval.cpp:155,169)."""
    import gzip
    text = gzip.open(os.path.join(golden_io.GOLDEN, "kat_primitives.txt.gz"), "rt").read()
    kat = os.path.join(ROOT, "oracle", "_ref", "ref_kat")
    if os.path.exists(kat):
        assert subprocess.run([kat], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=600, check=True).stdout.decode() == text
    L = text.splitlines()
    # (1-2x+3x^2 +/-0.75)*(3+2x-4x^2 +/-0.5) = 3-4x+x^2+14x^3-12x^4 +/-10.125
    i = [k for k, l in enumerate(L) if l.startswith("tm mul as 1")][1]
    r = L.index("r 5 %s" % float(10.125).hex().replace("0x1.4400000000000p+3", "0x1.44p+3"), i)
    assert [float.fromhex(l.split()[-1]) for l in L[r + 1:r + 6]] == [3.0, -4.0, 1.0, 14.0, -12.0]
    # intersection with roundoff: 1.0583333333333331 +/- 0.10833333333333339
    assert any(l.startswith("r 1 ") and float.fromhex(l.split()[2]) == 0.10833333333333339 for l in L)
    assert any(l.startswith("t 0 ") and float.fromhex(l.split()[-1]) == 1.0583333333333331 for l in L)
    # Interval(1)/Interval(3), pow(Interval(3,5),-1) is rec: [0.33333333333333331,0.33333333333333337]
    iv = [l.split() for l in L if l.startswith("iv div a 0x1p+0 0x1p+0 b 0x1.8p+1 0x1.8p+1")][0]
    assert (float.fromhex(iv[-2]), float.fromhex(iv[-1])) == (0.33333333333333331, 0.33333333333333337)
