"""Helpers shared by the parity tests: engine-format inputs -> oracle-format inputs, comparisons."""
import numpy as np

from oracle import fithic_oracle as O


def oracle_inputs(contacts, frags, st, biases=None):
    """fithic_b200 Contacts/Fragments/Biases/Settings -> arguments of oracle.run_pipeline."""
    c1 = (contacts.chrs & 0xffff).astype(np.int32)
    c2 = (contacts.chrs >> 16).astype(np.int32)
    oc = O.Contacts(c1, contacts.mid1.astype(np.int64), c2, contacts.mid2.astype(np.int64),
                    contacts.cnt.astype(np.int64), list(frags.chroms))
    res = st.resolution
    fchr, fmid = [], []
    for ci in range(len(frags.chroms)):
        n = int(frags.n_mappable[ci])
        if n == 0:
            continue
        if getattr(frags, "mids", None) is not None:
            mids = np.asarray(frags.mids[ci], dtype=np.int64)  # restriction fragments: the mid points themselves
        else:
            # synthetic fragments: n loci, the last one at max_mid, spaced by res
            mids = frags.max_mid[ci] - (n - 1 - np.arange(n, dtype=np.int64)) * res
        fchr.append(np.full(n, ci, dtype=np.int32))
        fmid.append(mids)
    fchr = np.concatenate(fchr) if fchr else np.zeros(0, np.int32)
    fmid = np.concatenate(fmid) if fmid else np.zeros(0, np.int64)
    fh = np.ones(len(fmid), dtype=np.int64)
    ost = O.Settings(resolution=res, noOfBins=st.noOfBins, mappThres=st.mappThres, distLowThres=st.distLowThres,
                     distUpThres=st.distUpThres, interOnly=st.interOnly, allReg=st.allReg,
                     biasLowerBound=st.biasLowerBound, biasUpperBound=st.biasUpperBound, noOfPasses=st.noOfPasses)
    ob = None
    if biases is not None:
        ob = {}
        for ci in range(len(biases.chr_off) - 1):
            lo, hi = int(biases.chr_off[ci]), int(biases.chr_off[ci + 1])
            m = biases.mids[lo:hi]
            ok = m >= 0
            if ok.any():
                ob[ci] = dict(zip(m[ok].astype(np.int64).tolist(), biases.values[lo:hi][ok].tolist()))
    return oc, fchr, fmid, fh, ost, ob


TINY = 1e-290


def rel_err(a, b):
    """max |a-b|/|b| over entries where b is finite and not tiny; the classes {0, NaN} must agree exactly, except that
    results below 1e-290 only have to be below 1e-290 on both sides (CUDA's exp() returns 0 below e^-744.4 where glibc
    still returns a denormal; the reference prints both as 0.000000e+00)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    tiny = np.abs(b) < TINY
    assert np.all(np.abs(a[tiny]) < TINY), "a result that should be < 1e-290 is not"
    assert np.array_equal((a == 0) & ~tiny, (b == 0) & ~tiny), "zero pattern differs"
    m = ~np.isnan(b) & ~tiny
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def compare_pass(r, o, tol=1e-6, check_ones=True):
    """r: engine result dict (host numpy p/q/expcc), o: oracle pass dict.  Returns dict of max relative errors."""
    assert r["N"] == o["N"], (r["N"], o["N"])
    assert r["T"] == o["T"], (r["T"], o["T"])
    assert r["observedInterAllCount"] == o["observedInterAllCount"]
    assert r["observedInterAllSum"] == o["observedInterAllSum"]
    assert r["observedIntraAllSum"] == o["observedIntraAllSum"]
    assert np.array_equal(r["dists"], o["dists"])
    assert np.array_equal(r["sums"], o["sums"])
    ob = o["bins"]
    rb = r["bins"]
    assert rb["n"] == len(ob)
    for i, b in enumerate(ob):
        assert (int(rb["lb"][i]), int(rb["ub"][i]), int(rb["sumcc"][i]), int(rb["pairs"][i])) == \
            (b["lb"], b["ub"], b["sumcc"], b["pairs"]), (i, b)
        assert float(rb["sumdist"][i]) == b["sumdist"], (i, rb["sumdist"][i], b["sumdist"])
    out = {}
    if o["splineX"] is not None:
        assert list(r["x"]) == list(o["x"]) and list(r["y"]) == list(o["y"])
        assert np.array_equal(np.asarray(r["splineX"]), np.asarray(o["splineX"]))
        out["table"] = rel_err(r["table"], o["newSplineY"])
        assert out["table"] <= 1e-12, out
    p, q, e = r["p"], r["q"], r["expcc"]
    if check_ones:
        assert np.array_equal(p == 1.0, o["p"] == 1.0), "p == 1.0 class differs"
    out["p"] = rel_err(p, o["p"])
    out["q"] = rel_err(q, o["q"])
    if "expcc" in o:  # the reference fixtures carry p and q only (ExpCC is printed with %f)
        out["expcc"] = rel_err(e, o["expcc"])
        assert out["expcc"] <= 1e-12, out
    assert out["p"] <= tol and out["q"] <= tol, out
    return out


# ---------------------------------------------------------------------------------------------------------------------
# golden fixtures (tests/golden/*.npz, generated from the unmodified reference by tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------------------------
import json  # noqa: E402
import os  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["intra_40kb", "intra_bias_LU_p2", "all_bias", "inter_only_bias", "intra_p3"]
# every k-th line of the reference's own bundled data sets (tests/golden/make_golden_real.py): real count distributions,
# real ICE biases, real (irregular) fragment lists
REAL_CASES = ["real_pfal_10kb", "real_hesc_40kb_bias"]
# restriction-fragment mode (-r 0) on the reference's bundled HindIII fragments
R0_CASES = ["real_hesc_refrags_r0", "real_mesc_refrags_r0_bias"]


def load_kat():
    with open(os.path.join(GOLDEN_DIR, "kat.json")) as f:
        return json.load(f)


def load_golden(name):
    """-> (Contacts, Fragments, Biases or None, Settings, list of per-pass reference dicts)."""
    from fithic_b200.engine import Biases, Contacts, Fragments, Settings
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    chroms = [str(c) for c in z["chroms"]]
    contacts = Contacts(z["mid1"], z["mid2"], z["cnt"], z["chrs"], chroms)
    frags = Fragments(chroms, z["frag_n"], z["frag_maxmid"])
    if "frag_mids" in z:  # restriction-fragment mode: the mid points themselves
        cuts = np.concatenate([[0], np.cumsum(z["frag_n"])]).astype(np.int64)
        frags.mids = [z["frag_mids"][cuts[i]:cuts[i + 1]] for i in range(len(chroms))]
    biases = None
    if "bias_values" in z:
        biases = Biases(z["bias_values"], z["bias_mids"], z["bias_chr_off"], int(z["res"]) == 0)
    st = Settings(resolution=int(z["res"]))
    flags = [str(f) for f in z["flags"]]
    i = 0
    while i < len(flags):
        f, v = flags[i], flags[i + 1]
        if f == "-b":
            st.noOfBins = int(v)
        elif f == "-p":
            st.noOfPasses = int(v)
        elif f == "-L":
            st.distLowThres = int(v)
        elif f == "-U":
            st.distUpThres = int(v)
        elif f == "-x":
            st.interOnly = v == "interOnly"
            st.allReg = v == "All"
        i += 2
    passes = []
    for k in range(1, int(z["npasses"]) + 1):
        pre = "p%d_" % k
        sc = z[pre + "scalars"]
        nb = len(z[pre + "bin_lb"])
        p7 = z[pre + "bin_pairs7"] if (pre + "bin_pairs7") in z else z[pre + "bin_pairs"]
        bins = [dict(lb=int(z[pre + "bin_lb"][i]), ub=int(z[pre + "bin_ub"][i]), pairs=int(z[pre + "bin_pairs"][i]),
                     pairs7=int(p7[i]), sumcc=int(z[pre + "bin_sumcc"][i]), sumdist=float(z[pre + "bin_sumdist"][i]))
                for i in range(nb)]
        has_spline = (pre + "splineX") in z
        passes.append(dict(N=int(z[pre + "N"]), T=int(z[pre + "T"]), observedInterAllCount=int(sc[0]),
                           observedInterAllSum=int(sc[1]), observedIntraAllSum=int(sc[2]),
                           possibleIntraInRangeCount=int(sc[3]), dists=z[pre + "dists"], sums=z[pre + "sums"], bins=bins,
                           x=sorted(z[pre + "x"].tolist()) if has_spline else z[pre + "x"].tolist(),
                           y=[v for _, v in sorted(zip(z[pre + "x"].tolist(), z[pre + "y"].tolist()))]
                           if has_spline else z[pre + "y"].tolist(),
                           splineX=z[pre + "splineX"] if has_spline else None,
                           newSplineY=z[pre + "newSplineY"] if has_spline else None, p=z[pre + "p"], q=z[pre + "q"],
                           outliersline=z[pre + "outliersline"], outliersdist=z[pre + "outliersdist"],
                           interChrProb=float(z[pre + "interChrProb"])))
    extra = dict(sig_head=[str(s) for s in z["sig_head"]], sig_nrows=int(z["sig_nrows"]),
                 bias_raw=z["bias_raw"] if "bias_raw" in z else None)
    return contacts, frags, biases, st, passes, extra


def cut_hist_numpy(p, p_cut0):
    """numpy model of fhc_bh_cut_hist: buckets = high 16 bits of the rankable p-values below p_cut0."""
    from fithic_b200 import _capi
    p = np.asarray(p, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ok = ~((p == 1.0) | np.isnan(p) | (p >= p_cut0))
    v = p[ok]
    b = np.where(v > 0, np.minimum(v.view(np.uint64) >> np.uint64(47), np.uint64(_capi.BH_CUT_BUCKETS - 1)), 0)
    return np.bincount(b.astype(np.int64), minlength=_capi.BH_CUT_BUCKETS).astype(np.uint64)


def merge_components_host(chr_rank, b1, b2, cc, q, conn, top_pct, neigh, sort_order):
    """fhc_host_merge_*: the per-entry code of csrc/merge.cu run serially on host arrays (what the CPU tests check)."""
    from fithic_b200 import _capi
    lib = _capi.load()
    n = len(b1)
    m = max(n, 1)
    a = dict(keys=np.zeros(m, np.uint64), order=np.zeros(m, np.uint32), label=np.zeros(m, np.int32), size=np.zeros(m, np.int32),
             first_line=np.zeros(m, np.uint32), box=np.zeros(4 * m, np.int32), sum_cc=np.zeros(m, np.int64),
             have=np.zeros(m, np.int64), ranked=np.zeros(m, np.uint32), keep=np.zeros(m, np.uint8))
    ins = [np.ascontiguousarray(chr_rank, np.int32), np.ascontiguousarray(b1, np.int32), np.ascontiguousarray(b2, np.int32),
           np.ascontiguousarray(cc, np.int64)]
    qq = np.ascontiguousarray(q, np.float64)
    p = _capi.dptr
    _capi.check(lib.fhc_host_merge_components(p(ins[0]), p(ins[1]), p(ins[2]), p(ins[3]), n, int(conn), p(a["keys"]),
                                              p(a["order"]), p(a["label"]), p(a["size"]), p(a["first_line"]), p(a["box"]),
                                              p(a["sum_cc"]), p(a["have"])))
    select = 0 < top_pct <= 100
    if select:
        _capi.check(lib.fhc_host_merge_select(p(a["keys"]), p(a["order"]), p(a["label"]), p(a["size"]), p(ins[3]), p(qq), n,
                                              int(top_pct), int(neigh), int(sort_order), p(a["ranked"]), p(a["keep"])))
    out = {k: v[:n] for k, v in a.items() if k != "box"}
    out["box"] = a["box"][:4 * n].reshape(-1, 4)
    if not select:
        del out["ranked"], out["keep"]
    return out
