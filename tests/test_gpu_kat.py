"""Known-answer tests of the PRIMITIVES on the device: every TaylorModel operator / function and Interval operation the
hot path uses, replayed through the C ABI (ab200_taylor_model_op, ab200_interval_op) against vectors produced by the
unmodified reference library (oracle/harness/ref_kat.cpp -> tests/golden/kat_primitives.txt.gz). The vectors contain the
literal cases of the reference's own unit tests (test/test_taylor_model.cpp:164-177,197-199,224-270,
test/test_interval.cpp:152-254) plus pseudo-random sparse models whose coefficients round. Bit-exact, structure included."""
import gzip
import os

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def load_kat():
    tms, ivs = [], []
    cur = None; model = None
    for line in gzip.open(os.path.join(golden_io.GOLDEN, "kat_primitives.txt.gz"), "rt"):
        t = line.split()
        if not t:
            continue
        if t[0] == "tm":
            cur = {"op": t[1], "as": int(t[3]), "k": int(t[5]), "flag": int(t[7]), "s": float.fromhex(t[9]),
                   "iv": (float.fromhex(t[11]), float.fromhex(t[13])), "x": None, "y": None, "r": None, "v": None}
            tms.append(cur)
        elif t[0] in ("x", "y", "r"):
            model = {"exps": [], "coefs": [], "error": float.fromhex(t[2])}
            cur[t[0]] = model
        elif t[0] == "t":
            model["exps"].append([int(e) for e in t[1:-1]]); model["coefs"].append(float.fromhex(t[-1]))
        elif t[0] == "v":
            cur["v"] = (float.fromhex(t[1]), float.fromhex(t[2]))
        elif t[0] == "iv":
            ivs.append({"op": t[1], "a": (float.fromhex(t[3]), float.fromhex(t[4])), "b": (float.fromhex(t[6]), float.fromhex(t[7])),
                        "m": int(t[9]), "r": (float.fromhex(t[11]), float.fromhex(t[12]))})
    return tms, ivs


def _bits(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64)).view(np.uint64)


@pytest.fixture(scope="module")
def ev(pkg):
    e = pkg.Evolver(open(os.path.join(golden_io.GOLDEN, "vanderpol.sys")).read(), settings={"arena_words": 262144})
    yield e
    e.close()


def test_taylor_model_primitives_bit_exact(ev):
    tms, _ = load_kat()
    assert len(tms) >= 225
    seen = set(); bad = []
    for i, c in enumerate(tms):
        x = (c["x"]["exps"], c["x"]["coefs"], c["x"]["error"]); y = (c["y"]["exps"], c["y"]["coefs"], c["y"]["error"])
        what = "#%d %s (as=%d, %d x %d terms, k=%d flag=%d)" % (i, c["op"], c["as"], len(x[1]), len(y[1]), c["k"], c["flag"])
        try:
            got = ev.taylor_model_op(c["op"], c["as"], x, y, k=c["k"], flag=c["flag"], s=c["s"], interval=c["iv"])
        except Exception as ex:
            bad.append(what + ": " + str(ex)); continue
        if c["op"] == "range":
            if not np.array_equal(_bits(got), _bits(c["v"])):
                bad.append(what + ": range %r vs %r" % (got, c["v"]))
        else:
            e, v, err = got
            if e.tolist() != c["r"]["exps"]:
                bad.append(what + ": structure (%d vs %d terms)" % (len(v), len(c["r"]["coefs"])))
            elif not np.array_equal(_bits(v), _bits(c["r"]["coefs"])):
                bad.append(what + ": coefficients")
            elif _bits([err])[0] != _bits([c["r"]["error"]])[0]:
                bad.append(what + ": error %r vs %r" % (float(err).hex(), float(c["r"]["error"]).hex()))
        seen.add(c["op"])
    assert not bad, "\n".join(bad)
    assert seen >= {"add", "sub", "mul", "addf", "mulf", "divf", "addi", "muli", "divi", "neg", "intersection", "split",
                    "antiderivative", "rec", "sqrt", "abs", "max", "range", "exp", "log", "sin", "cos"}


def test_literal_vectors_of_the_reference_unit_tests(ev):
    """A few of the reference's assertions spelled out (test/test_taylor_model.cpp:164-177,231-234,266-267)."""
    a = ([[0], [1], [2]], [1.0, -2.0, 3.0], 0.75); b = ([[0], [1], [2]], [3.0, 2.0, -4.0], 0.5)
    e, v, err = ev.taylor_model_op("mul", 1, a, b)
    assert e[:, 0].tolist() == [0, 1, 2, 3, 4] and v.tolist() == [3.0, -4.0, 1.0, 14.0, -12.0] and err == 10.125
    e, v, err = ev.taylor_model_op("divi", 1, a, None, interval=(0.25, 2.0))
    assert v.tolist() == [2.25, -4.5, 6.75] and err == 13.5
    e, v, err = ev.taylor_model_op("intersection", 1, ([[0]], [2. / 3], 0.5), ([[0]], [6. / 5], 0.25))
    assert v.tolist() == [1.0583333333333331] and err == 0.10833333333333339
    e, v, err = ev.taylor_model_op("antiderivative", 1, ([[2]], [2.0], 1.0), None, k=0)
    assert e.tolist() == [[3]] and v.tolist() == [0.66666666666666663] and err == 1.0000000000000002


def test_interval_primitives_bit_exact(ev):
    _, ivs = load_kat()
    assert len(ivs) >= 350 and {c["op"] for c in ivs} >= {"exp", "log", "sin", "cos"}
    out = ev.interval_ops([(c["op"], c["m"], c["a"], c["b"]) for c in ivs])
    for c, r in zip(ivs, out):
        assert r[2] == 0, c
        assert np.array_equal(_bits(r[:2]), _bits(c["r"])), "%s %r %r m=%d: %r vs %r" % (c["op"], c["a"], c["b"], c["m"], tuple(r[:2]), c["r"])
