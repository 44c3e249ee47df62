"""CPU-side parity of the algorithm itself: the kernel headers and the analyser mirror are compiled for the HOST (a
development build, tests/tools/hostchain.cpp, with the device lap replaced by a serial loop over the same
evolve_cell code) and replayed against the goldens of the compiled reference. This is test infrastructure only -- the
shipped library has no such path -- but it pins the host logic (paving, jump sets, laps) and the arithmetic of the
kernels bit for bit without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")
SRC = [os.path.join(ROOT, "tests", "tools", "hostchain.cpp"),
       os.path.join(ROOT, "release-1.0_b200", "csrc", "ab200_system.cpp"),
       os.path.join(ROOT, "release-1.0_b200", "csrc", "ab200_paving.cpp")]


def _build(name, extra):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no host compiler")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-frounding-math", "-w"] + extra + ["-x", "c++", "-o", exe] + SRC, timeout=600)
    return exe


@pytest.fixture(scope="module")
def hostchain_wide():
    """The same tool with two-word MultiIndex keys (systems with more than two state variables)."""
    return _build("hostchain_test_kw2", ["-DAB_KW=2"])


@pytest.fixture(scope="module")
def hostchain():
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no host compiler")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "hostchain_test")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-frounding-math", "-w", "-x", "c++", "-o", exe] + SRC, timeout=600)
    return exe


CASES = [
    ("bouncingball", "bouncingball_chain_acc5.txt", ["chain", "acc=5", "domain=-10.1,10.1,-0.1,2.1", "init=0.0,0.001,1.999,2.0"]),
    ("bouncingball", "bouncingball_chain_acc6.txt", ["chain", "acc=6", "domain=-10.1,10.1,-0.1,2.1", "init=-2.0,0.5,0.5,1.5"]),
    ("watertank", "watertank_reach_acc3.txt", ["upper", "acc=3", "T=0.5", "domain=0.0,1.0,4.5,9.0", "loc=2", "init=1.0,1.0,7.6,7.7",
                                               "loc=1", "init=0.0,0.0,5.7,5.8"]),
    ("vanderpol", "vanderpol_reach_acc4.txt", ["upper", "acc=4", "T=2.0", "domain=-3,3,-3,3", "init=-1.6,-1.3,0.3,0.8"]),
    ("watertank", "watertank_reach_acc3_T2.txt", ["upper", "acc=3", "T=2", "domain=0.0,1.0,4.5,9.0", "loc=2", "init=1.0,1.0,6.0,7.5",
                                                  "loc=1", "init=0.0,0.0,6.0,7.5"]),
    ("bouncingball", "bouncingball_backward_chain_acc5.txt", ["bchain", "acc=5", "domain=-10.1,10.1,-0.1,2.1", "init=2.0,4.0,0.0,0.3"]),
    ("bouncingball", "bouncingball_lower_reach_acc4.txt", ["lower", "acc=4", "T=1.5", "domain=-10.1,10.1,-0.1,2.1", "init=0.0,0.5,1.0,2.0"]),
    ("watertank", "watertank_lower_reach_acc3.txt", ["lower", "acc=3", "T=0.5", "domain=0.0,1.0,4.5,9.0", "loc=2", "init=1.0,1.0,7.6,7.7",
                                                     "loc=1", "init=0.0,0.0,5.7,5.8"]),
]


@pytest.mark.parametrize("system,golden,args", CASES)
def test_host_build_reproduces_reference_cell_sets(hostchain, system, golden, args):
    g = os.path.join(ROOT, "tests", "golden")
    out = subprocess.run([hostchain, os.path.join(g, system + ".sys"), os.path.join(g, golden)] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
    assert "cell sets identical (bit-exact boxes)" in out, out[-2000:]


# epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627): laps of LOWER_SEMANTICS evolution whose final TaylorSets
# start the next lap; the cells AND the epsilon vector must equal the reference's bit for bit. The watertank case evolves the
# system with its interval parameter at the midpoint and draws the pruning from glibc's rand() sequence.
EPS_CASES = [
    ("bouncingball", "bouncingball_epslower_acc4.txt", ["epslower", "acc=4", "domain=-10.1,10.1,-0.1,2.1", "init=0.0,0.001,1.999,2.0"]),
    ("watertank", "watertank_epslower_acc3_prune.txt", ["epslower", "acc=3", "prune=1", "subsys=" + os.path.join(ROOT, "tests", "golden", "watertank_midparams.sys"),
                                                        "domain=0.0,1.0,4.5,9.0", "loc=2", "init=1.0,1.0,6.0,7.5", "loc=1", "init=0.0,0.0,6.0,7.5"]),
]


@pytest.mark.parametrize("system,golden,args", EPS_CASES)
def test_host_build_epsilon_lower_chain_reach(hostchain, system, golden, args):
    g = os.path.join(ROOT, "tests", "golden")
    out = subprocess.run([hostchain, os.path.join(g, system + ".sys"), os.path.join(g, golden)] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
    assert "cell sets identical (bit-exact boxes)" in out and "differs" not in out, out[-2000:]


def test_host_build_wide_keys_epsilon_lower_chain_reach(hostchain_wide):
    """The enclosure hand-over with two-word MultiIndex keys (springs, five variables; three words per term in the records)."""
    g = os.path.join(ROOT, "tests", "golden")
    args = ["epslower", "acc=2", "prune=1", "lock=1.0", "domain=0.0,40.0,-2.0,2.0,-3.0,3.0,-1.0,3.0,0.0,4.0", "init=0,0,0,0,0,0,0,0,3,3"]
    out = subprocess.run([hostchain_wide, os.path.join(g, "springs.sys"), os.path.join(g, "springs_epslower_acc2_lock1.txt")] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
    assert "cell sets identical (bit-exact boxes)" in out and "differs" not in out and "laps 3" in out, out[-2000:]


# constraint-set initial sets: the golden carries the set's text (cons| lines); the tool discretises it
# (ab200_analyser_discretise_constraint) and runs the analysis from the cells / from the midpoints the set covers
CONS_CASES = [
    ("vanderpol", "vanderpol_cons_band_acc6.txt", ["upper", "acc=6", "T=0.02", "domain=-3,3,-3,3"]),
    ("bouncingball", "bouncingball_cons_disc_chain_acc5.txt", ["chain", "acc=5", "domain=-10.1,10.1,-0.1,2.1"]),
    ("vanderpol", "vanderpol_cons_disc_lower_reach_acc4.txt", ["lower", "acc=4", "T=1.5", "domain=-3,3,-3,3"]),
    ("vanderpol", "vanderpol_cons_disc_epslower_acc4.txt", ["epslower", "acc=4", "prune=1", "domain=-3,3,-3,3"]),
]


@pytest.mark.parametrize("system,golden,args", CONS_CASES)
def test_host_build_constraint_set_initial_sets(hostchain, system, golden, args):
    g = os.path.join(ROOT, "tests", "golden")
    out = subprocess.run([hostchain, os.path.join(g, system + ".sys"), os.path.join(g, golden)] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
    assert "cell sets identical (bit-exact boxes)" in out and "differs" not in out, out[-2000:]


def test_host_build_reproduces_the_reference_domain_error(hostchain_wide):
    """epsilon_lower_chain_reach of the boost converter: the reference throws "The lower reach is not inside the over-approximated
    domain" (src/reachability_analyser.cpp:596-599; the golden holds its message); the mirror returns AB200_E_DOMAIN with it."""
    g = os.path.join(ROOT, "tests", "golden")
    assert "The lower reach is not inside the over-approximated domain" in open(os.path.join(g, "boost_epslower_acc3_domain_error.txt")).read()
    args = ["epslower", "acc=3", "prune=1", "domain=0.0,0.001,0.0,3.0,0.0,6.0", "loc=1", "init=0.0,0.0,1.0,1.0,1.0,1.0"]
    out = subprocess.run([hostchain_wide, os.path.join(g, "boost.sys"), os.path.join(g, "boost_epslower_acc3_domain_error.txt")] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
    assert "rc=3 The lower reach is not inside the over-approximated domain" in out, out[-2000:]


WIDE_CASES = [
    ("springs", "springs_reach_acc3.txt.gz", ["upper", "acc=3", "T=40", "domain=0.0,40.0,-2.0,2.0,-3.0,3.0,-1.0,3.0,0.0,4.0",
                                              "init=0,0,0,0,0,0,0,0,3,3"]),
    ("boost", "boost_reach_acc4.txt", ["upper", "acc=4", "T=0.011", "domain=0.0,0.001,0.0,3.0,0.0,6.0", "loc=1", "init=0.0,0.0,1.0,1.0,1.0,1.0"]),
]


@pytest.mark.parametrize("system,golden,args", WIDE_CASES)
def test_host_build_wide_keys_reproduces_reference_cell_sets(hostchain_wide, system, golden, args, tmp_path):
    import gzip
    g = os.path.join(ROOT, "tests", "golden")
    path = os.path.join(g, golden)
    if golden.endswith(".gz"):
        path = str(tmp_path / golden[:-3])
        with open(path, "wb") as f:
            f.write(gzip.open(os.path.join(g, golden)).read())
    out = subprocess.run([hostchain_wide, os.path.join(g, system + ".sys"), path] + args,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900).stdout.decode()
    assert "cell sets identical (bit-exact boxes)" in out, out[-2000:]


def test_primitive_known_answers_on_the_host_build(tmp_path):
    """Every TaylorModel operator / function and Interval operation of the hot path, replayed on the CPU: the kernel headers
    compiled for the host (fenv rounding, the cooperative primitives under the 32-lane fiber emulator), through the same
    dispatch the device kernels use (ab200_kat.cuh), against the vectors the reference library produced
    (tests/golden/kat_primitives.txt.gz: 230 model cases incl. exp / log / sin / cos / rec / sqrt, 358 interval cases)."""
    import gzip
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no host compiler")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "hostkat_test")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-frounding-math", "-w", "-o", exe, os.path.join(ROOT, "tests", "tools", "hostkat.cpp")], timeout=600)
    vec = str(tmp_path / "kat.txt")
    with open(vec, "wb") as f:
        f.write(gzip.open(os.path.join(ROOT, "tests", "golden", "kat_primitives.txt.gz")).read())
    for lanes in ("32", "1"):
        out = subprocess.run([exe, vec, lanes], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
        assert "230 TaylorModel cases, 0 mismatching; 358 Interval cases, 0 mismatching" in out, out[-2000:]
