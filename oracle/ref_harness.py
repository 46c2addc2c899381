"""Run the UNMODIFIED reference (ay-lab/fithic) in this container and capture full-precision intermediates.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fithic_b200/) may import this file.  It needs
/root/reference, which exists only in the build container -- never on the GPU box -- so it is used solely by
tests/golden/make_golden.py (to generate the committed fixtures) and by container-only tests that are skipped when
/root/reference is absent.

The reference imports matplotlib/pylab at module scope (fithic/fithic.py:30-34), which are not installed here; they
are replaced by inert stubs (plots are out of scope and `-v` is never passed).  No reference file is modified.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FITHIC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fithic", "fithic.py"))


class _Stub(types.ModuleType):
    __all__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


def load_reference():
    """Import fithic.fithic from /root/reference with matplotlib/pylab stubbed (fithic/fithic.py:30-34)."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "pylab"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    tick = sys.modules["matplotlib.ticker"]
    for attr in ("ScalarFormatter", "FormatStrFormatter", "MaxNLocator"):
        if isinstance(tick, _Stub):
            setattr(tick, attr, lambda *a, **k: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fithic.fithic as F  # noqa: E402  (the reference, unmodified)
    return F


def run_reference(argv, quiet=True):
    """Run reference main() with `argv` (list of CLI flags) and capture intermediates of every pass.

    Returns a list with one dict per spline pass:
      mainDic (dist -> ΣCC), N, binStats rows, x, y, splineX, newSplineY, p, q, T, outliersline, outliersdist
    Capture is by rebinding module attributes; reference code itself is untouched.
    """
    F = load_reference()
    passes = []
    cur = {}

    orig_bh = F.myStats.benjamini_hochberg_correction
    orig_cp = F.calculateProbabilities
    orig_fs = F.fit_Spline
    orig_gf = F.generate_FragPairs
    orig_ri = F.read_Interactions

    def ri(*a, **k):
        out = orig_ri(*a, **k)
        cur["observedInterAllCount"], cur["observedInterAllSum"] = out[1], out[2]
        cur["observedIntraAllSum"], cur["N"] = out[3], out[4]
        return out

    def gf(*a, **k):
        out = orig_gf(*a, **k)
        (_, cur["noOfFrags"], cur["maxPossibleGenomicDist"], cur["possibleIntraInRangeCount"],
         cur["possibleInterAllCount"], cur["interChrProb"], cur["baselineIntraChrProb"]) = out
        return out

    def cp(mainDic, binStats, *a, **k):
        out = orig_cp(mainDic, binStats, *a, **k)
        cur["mainDic"] = {int(d): int(v[1]) for d, v in mainDic.items()}
        cur["bins"] = [dict(lb=int(b[0][0]), ub=int(b[0][1]), pairs=int(b[1]), sumcc=int(b[2]), sumdist=float(b[3]),
                            pairs7=int(b[7]), dists=[int(d) for d in b[6]]) for b in
                       (binStats[i] for i in range(len(binStats)))]
        cur["x"], cur["y"] = [float(v) for v in out[0]], [float(v) for v in out[1]]
        return out

    def bh(p_vals, T):
        q = orig_bh(p_vals, T)
        cur["p"], cur["q"], cur["T"] = list(p_vals), list(q), T
        return q

    def fs(*a, **k):
        out = orig_fs(*a, **k)
        cur["splineX"] = None if out[0] is None else [int(v) for v in out[0]]
        cur["newSplineY"] = None if out[1] is None else [float(v) for v in out[1]]
        cur["outliersline"], cur["outliersdist"] = [int(v) for v in out[3]], [int(v) for v in out[4]]
        passes.append(dict(cur))
        cur.clear()
        return out

    F.read_Interactions, F.generate_FragPairs, F.calculateProbabilities, F.fit_Spline = ri, gf, cp, fs
    F.myStats.benjamini_hochberg_correction = bh
    old_argv = sys.argv
    sys.argv = ["fithic"] + [str(a) for a in argv]
    try:
        with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
            F.main()
    finally:
        sys.argv = old_argv
        F.read_Interactions, F.generate_FragPairs, F.calculateProbabilities, F.fit_Spline = \
            orig_ri, orig_gf, orig_cp, orig_fs
        F.myStats.benjamini_hochberg_correction = orig_bh
    return passes
