"""CPU: the oracle (oracle/) pinned against fixtures captured from the UNMODIFIED reference (tests/golden/)."""
import numpy as np
import pytest

from oracle import fithic_oracle as O
from tests.util import GOLDEN_CASES, R0_CASES, load_golden, load_kat, oracle_inputs, rel_err


def ulps(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = ~np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return np.max(np.abs(a[m] - b[m]) / np.maximum(np.spacing(np.abs(b[m])), 5e-324)) if m.any() else 0.0


def test_c_cephes_bdtrc_against_scipy_vectors():
    """oracle/cephes_bdtrc.c against scipy.special.bdtrc known answers (third-party arithmetic)."""
    kat = load_kat()["bdtrc"]
    k = np.array([r[0] for r in kat], dtype=np.float64)
    N = np.array([r[1] for r in kat], dtype=np.int64)
    p = np.array([r[2] for r in kat], dtype=np.float64)
    want = np.array([r[3] for r in kat], dtype=np.float64)
    got = np.array([O.bdtrc([k[i]], int(N[i]), [p[i]])[0] for i in range(len(kat))])
    assert ulps(got, want) <= 4  # pow/log differ by an ulp or two between libm builds; typically 0


def test_bh_against_reference_vectors():
    for p, T, want in load_kat()["bh"]:
        got = O.benjamini_hochberg(np.array(p, dtype=np.float64), T)
        assert np.array_equal(got, np.array(want, dtype=np.float64), equal_nan=True)
        assert np.array_equal(np.array(O.benjamini_hochberg_loop(p, T), dtype=np.float64),
                              np.array(want, dtype=np.float64), equal_nan=True)


@pytest.mark.parametrize("name", GOLDEN_CASES + R0_CASES)
def test_oracle_pipeline_against_reference(name):
    contacts, frags, biases, st, ref, _ = load_golden(name)
    oc, fchr, fmid, fh, ost, ob = oracle_inputs(contacts, frags, st, biases)
    got = O.run_pipeline(oc, fchr, fmid, fh, ost, ob)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g["N"] == r["N"] and g["T"] == r["T"]
        assert g["observedInterAllCount"] == r["observedInterAllCount"]
        assert g["observedInterAllSum"] == r["observedInterAllSum"]
        assert g["observedIntraAllSum"] == r["observedIntraAllSum"]
        assert g["possibleIntraInRangeCount"] == r["possibleIntraInRangeCount"]
        assert np.array_equal(g["dists"], r["dists"]) and np.array_equal(g["sums"], r["sums"])
        assert len(g["bins"]) == len(r["bins"])
        for a, b in zip(g["bins"], r["bins"]):
            assert (a["lb"], a["ub"], a["pairs"], a["sumcc"]) == (b["lb"], b["ub"], b["pairs"], b["sumcc"])
            assert a["pairs7"] == b["pairs7"]  # differs from `pairs` only in restriction-fragment mode (:733-734)
            assert a["sumdist"] == b["sumdist"]
        assert list(g["x"]) == list(r["x"]) and list(g["y"]) == list(r["y"])
        if r["splineX"] is not None:
            assert np.array_equal(np.asarray(g["splineX"]), r["splineX"])
            assert np.array_equal(g["newSplineY"], r["newSplineY"])
        assert ulps(g["p"], r["p"]) <= 4
        assert rel_err(g["q"], r["q"]) <= 1e-15
        assert np.array_equal(np.asarray(g["outliersline"]), r["outliersline"])
        assert np.array_equal(np.asarray(g["outliersdist"]), r["outliersdist"])
