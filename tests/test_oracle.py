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


@pytest.mark.parametrize("name", ["vanderpol", "bouncingball", "watertank", "springs", "boost"])
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
