"""CPU-only tests of the native host stage between K1 and K3 (csrc/hoststage.cu, csrc/fitpack_host.cuh): the spline fit
against scipy's FITPACK bit for bit (scipy is the checker here, the product path does not import it), the fused stage
against the stage-by-stage entry points and the oracle."""
import ctypes
import warnings

import numpy as np
import pytest

from fithic_b200 import _capi, synth
from fithic_b200._capi import check, dptr
from fithic_b200.engine import Settings, calculate_probabilities, fit_spline, frag_pairs, make_bins
from oracle import fithic_oracle as O
from tests.util import GOLDEN_CASES, GOLDEN_DIR, R0_CASES, REAL_CASES


def curfit(lib, x, y, s):
    m = len(x)
    x, y = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
    t, c = np.zeros(m + 4), np.zeros(m + 4)
    n, ier, calls, fp = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
    check(lib.fhc_host_curfit(dptr(x), dptr(y), m, float(s), dptr(t), dptr(c), ctypes.byref(n), ctypes.byref(fp),
                              ctypes.byref(ier), ctypes.byref(calls)))
    return t[:n.value], c[:n.value], ier.value, calls.value


def scipy_tck(x, y, s):
    from scipy.interpolate import UnivariateSpline
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return UnivariateSpline(x, y, s=s)._eval_args


def test_curfit_equals_scipy_on_the_reference_fixtures(lib):
    """(x, y) of every pass of every fixture captured from the unmodified reference: knots and coefficients of
    UnivariateSpline(x, y, s=min(y)**2) (fithic/fithic.py:951) bit for bit."""
    import os
    seen = 0
    for name in GOLDEN_CASES + REAL_CASES + R0_CASES:
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        for p in range(1, int(z["npasses"]) + 1):
            if "p%d_splineX" % p not in z:
                continue
            x, y = z["p%d_x" % p], z["p%d_y" % p]
            o = np.argsort(x, kind="stable")
            xs, ys = x[o], y[o]
            t, c, k = scipy_tck(xs, ys, min(ys) ** 2)
            t2, c2, ier, calls = curfit(lib, xs, ys, min(ys) ** 2)
            assert np.array_equal(t, t2) and np.array_equal(c, c2), (name, p, len(t), len(t2), ier, calls)
            seen += 1
    assert seen >= 8


@pytest.mark.parametrize("m", [4, 5, 8, 17, 18, 30, 64, 100, 101, 200])
def test_curfit_equals_scipy_on_random_curves(lib, m):
    """Power-law curves with noise (smooth ones end with the polynomial or a few knots, noisy ones run into the storage
    limit of the first call and are continued with nest = m + 4 like UnivariateSpline._reset_nest), plus s = 0."""
    routes = set()
    for noise in (0.0, 1e-4, 0.01, 0.1, 0.3):
        for seed in range(4):
            rng = np.random.default_rng(seed * 1000 + m)
            x = np.sort(rng.uniform(5e3, 2e8, m))
            y = np.abs(1e-3 * (x / 5e3) ** -1.1 * (1 + noise * rng.standard_normal(m)))
            s = min(y) ** 2
            t, c, k = scipy_tck(x, y, s)
            t2, c2, ier, calls = curfit(lib, x, y, s)
            assert np.array_equal(t, t2) and np.array_equal(c, c2), (m, noise, seed, len(t), len(t2), ier, calls)
            routes.add((calls, ier))
    x = np.sort(np.random.default_rng(m).uniform(5e3, 2e8, m))
    y = 1e-3 * (x / 5e3) ** -1.1
    y[m // 2] = 0.0  # s = min(y)^2 = 0: the interpolating spline
    t, c, k = scipy_tck(x, y, 0.0)
    t2, c2, ier, calls = curfit(lib, x, y, 0.0)
    assert np.array_equal(t, t2) and np.array_equal(c, c2) and ier == -1
    assert len(routes) >= 1


def test_fit_spline_is_native_and_refuses_what_the_reference_refuses(lib, capsys):
    xs, ys, (t, c, k) = fit_spline([3.0, 1.0, 2.0, 4.0, 5.0], [0.1, 0.5, 0.3, 0.05, 0.01])
    assert xs == [1.0, 2.0, 3.0, 4.0, 5.0] and ys == [0.5, 0.3, 0.1, 0.05, 0.01] and k == 3
    tt, cc, _ = scipy_tck(xs, ys, min(ys) ** 2)
    assert np.array_equal(t, tt) and np.array_equal(c, cc)
    with pytest.raises(SystemExit) as e:  # fithic/fithic.py:940-945
        fit_spline([1.0, 1.0, 2.0, 3.0, 4.0], [0.5, 0.4, 0.3, 0.2, 0.1])
    assert e.value.code == 2
    assert "Distances do not decrease across bins" in capsys.readouterr().out


def _stage_io(lib, hist, scal, present, res, nbins, frags, L, U, want_spline, nthreads, dec=None, lbeta_cap=1 << 14):
    D = len(hist)
    nw = (D + 31) // 32
    k1 = np.zeros(D + _capi.N_SCALARS + (nw + 1) // 2, dtype=np.uint64)
    k1[:D] = hist
    k1[D:D + _capi.N_SCALARS] = scal
    if present is not None:
        k1[D + _capi.N_SCALARS:].view(np.uint32)[:nw] = present
    order = [i for i in sorted(range(len(frags.chroms)), key=lambda i: frags.chroms[i]) if frags.n_mappable[i] > 0]
    keep = dict(k1=k1, chr_n=np.ascontiguousarray(frags.n_mappable[order], dtype=np.int64),
                chr_mm=np.ascontiguousarray(frags.max_mid[order], dtype=np.int64),
                i64=np.zeros(3 * D + 4 * nbins, dtype=np.int64), f64=np.zeros(2 * D + 5 * nbins + 2 * (nbins + 4)),
                lb=[np.full(lbeta_cap, -7.0), np.full(lbeta_cap, -7.0)], dec=dec)
    io = _capi.StageIO()
    io.k1buf, io.D, io.grid, io.noOfBins, io.L, io.U = k1.ctypes.data, D, res, nbins, L, U
    io.chr_n, io.chr_maxmid, io.nchr = keep["chr_n"].ctypes.data, keep["chr_mm"].ctypes.data, len(order)
    io.want_spline, io.nthreads = want_spline, nthreads
    b = keep["i64"].ctypes.data
    io.dists, io.sums, io.splineX = b, b + 8 * D, b + 16 * D
    b += 24 * D
    io.bin_lb, io.bin_ub, io.bin_sumcc, io.bin_pairs = b, b + 8 * nbins, b + 16 * nbins, b + 24 * nbins
    b = keep["f64"].ctypes.data
    io.table, io.lut = b, b + 8 * D
    b += 16 * D
    io.bin_sumdist, io.x_bins, io.y_bins, io.xs, io.ys = (b + 8 * nbins * k for k in range(5))
    b += 40 * nbins
    io.t, io.c = b, b + 8 * (nbins + 4)
    for w in (0, 1):
        io.lbeta_tab[w], io.lbeta_cap[w] = keep["lb"][w].ctypes.data, lbeta_cap
    if dec is not None:
        io.dec = dec.ctypes.data
    return io, keep


def _views(io, keep, nbins):
    D, i64, f64 = int(io.D), keep["i64"], keep["f64"]
    n, ns, m, nt = int(io.nb), int(io.nseen), int(io.m), int(io.nt)
    o = 3 * D
    v = dict(dists=i64[:ns], sums=i64[D:D + ns], splineX=i64[2 * D:2 * D + m], lb=i64[o:o + n], ub=i64[o + nbins:o + nbins + n],
             sumcc=i64[o + 2 * nbins:o + 2 * nbins + n], pairs=i64[o + 3 * nbins:o + 3 * nbins + n], table=f64[:m],
             lut=f64[D:2 * D])
    o = 2 * D
    for k, name in enumerate(("sumdist", "x_bins", "y_bins", "xs", "ys")):
        v[name] = f64[o + k * nbins:o + k * nbins + n]
    o = 2 * D + 5 * nbins
    v["t"], v["c"] = f64[o:o + nt], f64[o + nbins + 4:o + nbins + 4 + nt]
    return v


STAGE_CASES = [
    # res, pairs, chroms, nbins, L, U, threads
    (40000, 200_000, ["chr1"], 100, 0, -1, 1),
    (40000, 200_000, ["chr1"], 50, 80000, 5000000, 4),
    (100000, 300_000, None, 100, 0, -1, 3),
    (10000, 150_000, ["chr20", "chr21", "chr22"], 200, 0, -1, 8),
]


@pytest.mark.parametrize("case", STAGE_CASES, ids=["%d_%d_%d" % (c[0], c[3], c[6]) for c in STAGE_CASES])
def test_host_stage_equals_the_stagewise_path(lib, case):
    """Histogram in, lookup table out: bins / pairs / sums of distances / x / y / totals / knots equal to the separate
    entry points (themselves pinned to the reference's fixtures in test_host.py) and scipy; table against the oracle's
    antitonic regression; lookup table against the clamp + bisect of fithic/fithic.py:1066-1068; lbeta table against
    fhc_host_lbeta_table."""
    res, npairs, chroms, nbins, L, U, threads = case
    contacts, frags, _, _ = synth.make_intra(npairs, res, seed=77 + nbins, chroms=chroms, mean_count=4.0)
    d = np.abs(contacts.mid1.astype(np.int64) - contacts.mid2)
    D = int(max(d.max(), frags.max_mid.max()) // res + 2)
    inr = (d >= L) & ((d <= U) if U >= 0 else True)
    hist = np.bincount(d[inr] // res, weights=contacts.cnt[inr], minlength=D).astype(np.int64)
    N = int(hist.sum())
    scal = np.zeros(_capi.N_SCALARS, dtype=np.uint64)
    scal[_capi.S_INTRA_INRANGE_SUM] = N
    scal[_capi.S_MAX_COUNT] = int(contacts.cnt.max())
    scal[_capi.S_INTER_ALL_SUM] = 12345
    rng = np.random.default_rng(5)
    for use_dec in (False, True):
        dec = rng.integers(0, 5, nbins).astype(np.int64) if use_dec else None
        io, keep = _stage_io(lib, hist, scal, None, res, nbins, frags, L, U, 1, threads, dec)
        if use_dec:  # pass >= 2: bins first, then the decrements, then the rest
            check(lib.fhc_host_stage(ctypes.byref(io), 1))
            check(lib.fhc_host_stage(ctypes.byref(io), 6))
        else:
            check(lib.fhc_host_stage(ctypes.byref(io), 7))
        assert io.status == 0
        v = _views(io, keep, nbins)
        seen = np.nonzero(hist)[0]
        st = Settings(resolution=res, noOfBins=nbins, distLowThres=L, distUpThres=float("inf") if U < 0 else U)
        bins = make_bins(lib, seen * res, hist[seen], nbins, N)
        fp = frag_pairs(lib, frags, st, bins, dec)
        x, y = calculate_probabilities(bins, N)
        assert np.array_equal(v["dists"], seen * res) and np.array_equal(v["sums"], hist[seen])
        for k in ("lb", "ub", "sumcc", "pairs"):
            assert np.array_equal(v[k], bins[k]), k
        assert np.array_equal(v["sumdist"], bins["sumdist"])  # the same double, term by term in the reference's order
        assert v["x_bins"].tolist() == x and v["y_bins"].tolist() == y
        assert [int(t) for t in io.totals] == [fp["possibleIntraInRangeCount"], int(fp["possibleIntraAllCount"] * 2),
                                               int(fp["possibleInterAllCount"] * 2), fp["noOfFrags"]]
        xs, ys, (t, c, k) = fit_spline(x, y)
        assert v["xs"].tolist() == xs and v["ys"].tolist() == ys
        tt, cc, _ = scipy_tck(xs, ys, min(ys) ** 2)
        assert np.array_equal(v["t"], tt) and np.array_equal(v["c"], cc)
        dists = seen * res
        sx = dists[(dists >= min(xs)) & (dists <= max(xs))]
        assert np.array_equal(v["splineX"], sx)
        osx, want, _, _, _ = O.fit_spline_table(x, y, dists, True)  # scipy spline + sklearn isotonic, as the reference
        assert osx == sx.tolist()
        assert np.max(np.abs(v["table"] - want) / np.abs(want)) <= 1e-13
        dl = np.clip(np.arange(D) * res, min(xs), max(xs))
        idx = np.minimum(np.searchsorted(sx, dl, side="left"), len(sx) - 1)
        assert np.array_equal(v["lut"], v["table"][idx])
        for w, Nw in ((0, N), (1, 12345)):
            ntab = int(io.lbeta_ntab[w])
            assert ntab == min(max(int(contacts.cnt.max()), 1), min(Nw, (1 << 22) - 1)) + 1
            ref = np.zeros(ntab)
            check(lib.fhc_host_lbeta_table(Nw, dptr(ref), ntab, 1))
            assert np.array_equal(keep["lb"][w][:ntab], ref, equal_nan=True)
            assert np.all(keep["lb"][w][ntab:] == -7.0)


def test_host_stage_sees_distances_whose_counts_sum_to_zero(lib):
    """read_Interactions keeps a distance in its dictionary even when its counts sum to zero (fithic/fithic.py:434-436):
    the stage consults K1's `present` bitmap when (and only when) K1 counted lines with a count <= 0."""
    res, nbins = 10000, 10
    frags = synth.fragments_for(["chr21"], np.array([46709983]), res)
    D = int(frags.max_mid.max() // res + 2)
    hist = np.zeros(D, dtype=np.int64)
    hist[[0, 1, 2, 5, 9]] = [50, 30, 10, 6, 4]
    present = np.zeros((D + 31) // 32, dtype=np.uint32)
    present[0] = (1 << 3) | (1 << 7)  # distances 3 and 7 were seen with counts of zero
    scal = np.zeros(_capi.N_SCALARS, dtype=np.uint64)
    scal[_capi.S_INTRA_INRANGE_SUM] = 100
    scal[_capi.S_MAX_COUNT] = 7
    for nonpos, want in ((0, [0, 1, 2, 5, 9]), (2, [0, 1, 2, 3, 5, 7, 9])):
        scal[_capi.S_NONPOS_LINES] = nonpos
        io, keep = _stage_io(lib, hist, scal, present, res, nbins, frags, 0, -1, 1, 1)
        check(lib.fhc_host_stage(ctypes.byref(io), 3))
        v = _views(io, keep, nbins)
        assert v["dists"].tolist() == [k * res for k in want]
        bins = make_bins(lib, np.array(want) * res, hist[want], nbins, 100)
        assert np.array_equal(v["ub"], bins["ub"])


def test_host_stage_reports_instead_of_fitting_nonsense(lib):
    res, nbins = 10000, 3
    frags = synth.fragments_for(["chr21"], np.array([46709983]), res)
    D = int(frags.max_mid.max() // res + 2)
    hist = np.zeros(D, dtype=np.int64)
    hist[:6] = [50, 30, 10, 6, 3, 1]
    scal = np.zeros(_capi.N_SCALARS, dtype=np.uint64)
    scal[_capi.S_INTRA_INRANGE_SUM] = 100
    scal[_capi.S_MAX_COUNT] = 1 << 15
    io, keep = _stage_io(lib, hist, scal, None, res, nbins, frags, 0, -1, 1, 2, lbeta_cap=64)
    check(lib.fhc_host_stage(ctypes.byref(io), 7))
    assert io.status == 4 and io.nb == 3  # three bins: scipy refuses a cubic fit, so does the stage
    assert io.lbeta_ntab[0] == 101  # min(max_count, N) + 1 does not fit 64 entries: the caller enlarges and asks again
    io2, keep2 = _stage_io(lib, hist, scal, None, res, nbins, frags, 0, -1, 0, 2, lbeta_cap=128)
    check(lib.fhc_host_stage(ctypes.byref(io2), 3))
    check(lib.fhc_host_stage(ctypes.byref(io2), 8))
    ref = np.zeros(101)
    check(lib.fhc_host_lbeta_table(100, dptr(ref), 101, 1))
    assert np.array_equal(keep2["lb"][0][:101], ref, equal_nan=True)


def test_pool_runs_every_job_exactly_once(lib):
    n = ctypes.c_int32()
    for threads in (1, 2, 5):
        for jobs in (1, 3, 40):
            ms = lib.fhc_host_pool_selftest(threads, jobs, 5, ctypes.byref(n))
            assert ms >= 0 and 1 <= n.value <= 64  # (the pool is shared: workers started by earlier calls join in)


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_host_stage_bins_on_ragged_histograms(lib, threads):
    """Phase 1 of fhc_host_stage (observed distances compacted in chunks on the pool, bins by bisection on the running sum of
    the counts) against fhc_host_make_bins' entry-by-entry loop on numpy-compacted input: histograms with long empty
    stretches, single distances that fill a bin on their own, distances seen only through the bitmap, lengths around the
    chunk size, and the tail the last closed bin leaves out."""
    frags = synth.fragments_for(["chrA"], np.array([50_000_000]), 1000)
    rng = np.random.default_rng(11 + threads)
    for D in (1, 7, 4095, 4096, 4097, 12289, 30011):
        for fill in (1.0, 0.3, 0.01):
            hist = rng.integers(1, 1000, D).astype(np.uint64)
            hist[rng.random(D) >= fill] = 0
            if D > 100:
                hist[rng.integers(0, D, 3)] = 500_000  # more than a bin's share in one distance
            if hist.sum() == 0:
                hist[0] = 5
            nw = (D + 31) // 32
            bits = (rng.random(D) < 0.05) & (hist == 0)
            present = np.packbits(bits, bitorder="little")
            present = np.concatenate([present, np.zeros(4 * nw - len(present), np.uint8)]).view(np.uint32)
            for use_bitmap in (False, True):
                scal = np.zeros(_capi.N_SCALARS, dtype=np.uint64)
                scal[_capi.S_INTRA_INRANGE_SUM] = int(hist.sum())
                scal[_capi.S_NONPOS_LINES] = 1 if use_bitmap else 0
                for nbins in (1, 10, 100):
                    io, keep = _stage_io(lib, hist, scal, present if use_bitmap else None, 1000, nbins, frags, 0, -1, 0, threads)
                    seen = np.nonzero((hist != 0) | (bits if use_bitmap else False))[0]
                    dists = (seen * 1000).astype(np.int64)
                    sums = hist[seen].astype(np.int64)
                    try:
                        want = make_bins(lib, dists, sums, nbins, int(hist.sum()))
                    except _capi.FithicB200Error:
                        # zero-count distances behind the last count: every one of them would close a bin of its own
                        # (desired == 0); both refuse to write past noOfBins bins
                        with pytest.raises(_capi.FithicB200Error, match="more than noOfBins"):
                            check(lib.fhc_host_stage(ctypes.byref(io), 1))
                        continue
                    check(lib.fhc_host_stage(ctypes.byref(io), 1))
                    v = _views(io, keep, nbins)
                    assert np.array_equal(v["dists"], dists) and np.array_equal(v["sums"], sums)
                    assert int(io.nb) == want["n"]
                    for k in ("lb", "ub", "sumcc"):
                        assert np.array_equal(v[k], want[k][:want["n"]]), (D, fill, nbins, k)
