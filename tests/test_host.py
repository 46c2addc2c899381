"""CPU-only tests: the C-ABI library loads and exports what include/fithic_b200.h declares, the host-side stages are bit
exact against the oracle, the table arithmetic (host build of the device source) matches libm / cephes, text I/O."""
import ctypes
import gzip
import math
import os
import re

import numpy as np
import pytest

from fithic_b200 import _capi, synth
from fithic_b200 import io as fio
from fithic_b200.engine import Settings, calculate_probabilities, fit_spline, frag_pairs, make_bins
from fithic_b200._capi import check, dptr  # noqa: E402
from oracle import fithic_oracle as O
from tests.util import GOLDEN_CASES, REAL_CASES, cut_hist_numpy, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "fithic_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fhc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    raw = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "library does not export %s" % name
    assert declared == set(_capi._SIGNATURES), declared ^ set(_capi._SIGNATURES)
    assert lib.fhc_abi_version() == _capi.FHC_ABI_VERSION


def test_errors_are_reported_not_swallowed(lib):
    rc = lib.fhc_host_make_bins(None, None, 5, 10, 100, None, None, None)
    assert rc == _capi.FHC_E_INVALID
    assert b"fhc_host_make_bins" in lib.fhc_last_error()
    with pytest.raises(_capi.FithicB200Error):
        _capi.check(rc)
    # device entry points validate their arguments before touching the GPU
    assert lib.fhc_pvalues(7, None, None, None, None, None, None, 0, 1, None, None, None, 0, 0, 10, 0, -1, None, 0, 1, 1, 0.0, 0.5, 2.0,
                           None, 0, None, 0, None, 0, 0.0, None, None, None, None, None, None, 0, None) == _capi.FHC_E_INVALID
    assert lib.fhc_lbeta_table(1 << 31, ctypes.c_void_p(16), 4, None) == _capi.FHC_E_RANGE


def test_no_cpu_fallback_without_cuda():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from fithic_b200.engine import Engine, Fragments
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(Settings(resolution=10000), Fragments([], np.zeros(0, np.int64), np.zeros(0, np.int64)))


def test_log_cr_is_libm_log(lib):
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.integers(1, 1 << 31, 20000).astype(np.float64), np.arange(1, 3000, dtype=np.float64),
                         (1 << 31) - np.arange(1, 2000, dtype=np.float64)])
    bad = [x for x in xs if lib.fhc_host_log_cr(float(x)) != math.log(x)]
    assert not bad, bad[:5]


def test_lbeta_is_cephes_lbeta(lib):
    ol = O._lib()
    ol.oracle_lbeta.restype = ctypes.c_double
    ol.oracle_lbeta.argtypes = [ctypes.c_double, ctypes.c_double]
    rng = np.random.default_rng(4)
    for N in (172, 1000, 4219169, 10 ** 8, 9 * 10 ** 8, (1 << 31) - 1):
        cs = list(range(1, 200)) + [int(c) for c in rng.integers(1, min(N, 3 * 10 ** 6), 1500)]
        for c in cs:
            if c > N:
                continue
            assert lib.fhc_host_lbeta(float(c), float(N - c + 1)) == ol.oracle_lbeta(float(c), float(N - c + 1)), (N, c)
    # N + 1 <= MAXGAM: cephes divides Gamma values; the lgam form used here agrees to ~1e-14
    for c in range(1, 100):
        a, b = lib.fhc_host_lbeta(float(c), float(100 - c + 1)), ol.oracle_lbeta(float(c), float(100 - c + 1))
        assert abs(a - b) <= 1e-13 * abs(b)


@pytest.mark.parametrize("name", GOLDEN_CASES + REAL_CASES)
def test_host_stages_bit_exact_against_reference_fixture(lib, name):
    """make_bins / frag_pairs / calculate_probabilities / fit_spline on the reference's own histogram."""
    contacts, frags, biases, st, ref, _ = load_golden(name)
    r = ref[0]
    bins = make_bins(lib, r["dists"], r["sums"], st.noOfBins, r["N"])
    fp = frag_pairs(lib, frags, st, bins)
    assert bins["n"] == len(r["bins"])
    for i, b in enumerate(r["bins"]):
        assert (int(bins["lb"][i]), int(bins["ub"][i]), int(bins["sumcc"][i]), int(bins["pairs"][i])) == \
            (b["lb"], b["ub"], b["sumcc"], b["pairs"])
        assert float(bins["sumdist"][i]) == b["sumdist"]
    assert fp["possibleIntraInRangeCount"] == r["possibleIntraInRangeCount"]
    x, y = calculate_probabilities(bins, r["N"])
    if r["splineX"] is not None:
        xs, ys, tck = fit_spline(x, y)
        assert xs == list(r["x"]) and ys == list(r["y"])


def test_make_bins_known_answers(lib):
    """SURVEY.md Appendix B (outputs of the unmodified makeBinsFromInteractions)."""
    def run(d, nb, N):
        dists = np.array(sorted(d), dtype=np.int64)
        sums = np.array([d[k] for k in sorted(d)], dtype=np.int64)
        b = make_bins(lib, dists, sums, nb, N)
        return [((int(b["lb"][i]), int(b["ub"][i])), int(b["sumcc"][i])) for i in range(b["n"])]
    assert run({0: 50, 10: 30, 20: 5, 30: 5, 40: 5, 50: 3, 60: 1, 70: 1}, 4, 100) == \
        [((0, 0), 50), ((1, 10), 30), ((11, 30), 10), ((31, 70), 10)]
    assert run({10 * k: 10 for k in range(10)}, 3, 100) == [((0, 30), 40), ((31, 60), 30), ((61, 90), 30)]
    assert run({0: 1, 10: 1, 20: 1}, 5, 3) == [((0, 0), 1), ((1, 10), 1), ((11, 20), 1)]
    assert run({0: 7, 10: 0, 20: 3, 30: 0}, 2, 10) == [((0, 0), 7), ((1, 20), 3)]  # distance 30 is dropped
    assert run({}, 5, 0) == []


def test_frag_pairs_quirks(lib):
    """x2 in-range count, negative npairs with unmappable loci, L/U window."""
    from fithic_b200.engine import Fragments
    st = Settings(resolution=10)
    frags = Fragments(["b", "a"], np.array([3, 5], dtype=np.int64), np.array([45, 65], dtype=np.int64))
    bins = dict(n=2, lb=np.array([0, 11], np.int64), ub=np.array([10, 40], np.int64), sumcc=np.array([1, 1], np.int64))
    fp = frag_pairs(lib, frags, st, bins)
    # chr a: n=5, max mid 65 -> dists 0..60: npairs 5,4,3,2,1,0,-1 ; chr b: n=3, max mid 45 -> dists 0..40: 3,2,1,0,-1
    # bin 0 = [0, 10], bin 1 = [11, 40] and everything beyond (clamped to the last bin)
    assert list(bins["pairs"]) == [(5 + 4) + (3 + 2), (3 + 2 + 1 + 0 - 1) + (1 + 0 - 1)]
    assert fp["possibleIntraInRangeCount"] == 2 * ((5 + 4 + 3 + 2 + 1 + 0 - 1) + (3 + 2 + 1 + 0 - 1))
    # the same through the oracle's sequential restatement
    obins = [dict(lb=0, ub=10, pairs=0, sumcc=1, sumdist=0.0, pairs7=0), dict(lb=11, ub=40, pairs=0, sumcc=1, sumdist=0.0,
                                                                              pairs7=0)]
    fchr = np.array([0] * 3 + [1] * 5)
    fmid = np.array([25, 35, 45, 25, 35, 45, 55, 65])
    out = O.generate_frag_pairs(fchr, fmid, np.ones(8, dtype=np.int64), ["b", "a"], O.Settings(resolution=10), obins, 0)
    assert [b["pairs"] for b in obins] == list(bins["pairs"])
    assert [b["sumdist"] for b in obins] == list(bins["sumdist"])
    assert out[3] == fp["possibleIntraInRangeCount"]
    assert fp["noOfFrags"] == 8 and fp["possibleInterAllCount"] == 15.0


def test_bias_quantiles_are_scipys_mquantiles():
    """read_biases logs the 5th / 50th / 95th quantile of the bias values with scipy.stats.mstats.mquantiles
    (fithic/fithic.py:812-816); io._mquantiles restates it (the command line then never imports scipy.stats): same bits."""
    from scipy.stats.mstats import mquantiles
    from fithic_b200.io import _mquantiles
    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 4, 5, 7, 10, 19, 20, 21, 100, 1001, 50_000]:
        for rep in range(20):
            a = np.exp(rng.normal(0, 0.5, n))
            if rep % 3 == 0:
                a = np.round(a, 1)  # ties
            want = np.ma.getdata(mquantiles(a, prob=[0.05, 0.5, 0.95])).astype(np.float64)
            got = np.asarray(_mquantiles(a, (0.05, 0.5, 0.95)), dtype=np.float64)
            assert [float(v).hex() for v in want] == [float(v).hex() for v in got], (n, rep)


def test_io_round_trip_and_bias_semantics(tmp_path):
    contacts, frags, biases, raw = synth.make_intra(5000, 100000, 5, chroms=["chr21", "chr22"], with_bias=True,
                                                    inter_fraction=0.2)
    cpath, fpath, bpath = synth.write_inputs(str(tmp_path), contacts, frags, 100000, raw, biases)
    c2 = fio.read_contacts(cpath)
    cp = fio.read_contacts_pandas(cpath)  # the native reader against pandas
    assert c2.chroms == cp.chroms
    for k in ("mid1", "mid2", "cnt", "chrs"):
        assert np.array_equal(getattr(c2, k), getattr(cp, k)), k
    names = c2.chroms
    m = {n: contacts.chroms.index(n) for n in names}
    remap = np.array([m[n] for n in names])
    assert np.array_equal(c2.mid1, contacts.mid1) and np.array_equal(c2.cnt, contacts.cnt)
    assert np.array_equal(remap[c2.chrs & 0xffff], contacts.chrs & 0xffff)
    chroms = list(c2.chroms)
    f2 = fio.read_fragments(fpath, chroms, 1)
    for i, n in enumerate(f2.chroms):
        j = frags.chroms.index(n)
        assert f2.n_mappable[i] == frags.n_mappable[j] and f2.max_mid[i] == frags.max_mid[j]
    b2, log = fio.read_biases(bpath, chroms, 100000, 0.5, 2.0)
    for ci, n in enumerate(chroms):
        j = contacts.chroms.index(n)
        lo, hi = int(biases.chr_off[j]), int(biases.chr_off[j + 1])
        a = b2.values[int(b2.chr_off[ci]):int(b2.chr_off[ci + 1])]
        assert np.array_equal(a, biases.values[lo:hi])
    assert any("discarded" in line for line in log)
    # count truncation toward zero and first-occurrence-wins for biases (fithic/fithic.py:415, :831-832)
    p = tmp_path / "t.gz"
    with gzip.open(p, "wt") as f:
        f.write("c1\t5\tc1\t15\t2.9\nc1\t5\tc2\t25\t-1.5\n")
    c = fio.read_contacts(str(p))
    assert list(c.cnt) == [2, -1]
    with gzip.open(p, "wt") as f:
        f.write("c1  5 \t c1 15   7e2\n\nc1\t5\tc2\t25\t+3\n")  # mixed whitespace, blank line, exponent, sign
    c = fio.read_contacts(str(p))
    assert list(c.cnt) == [700, 3] and c.chroms == ["c1", "c2"] and list(c.mid2) == [15, 25]
    for bad in ("c1\t5\tc1\t15\n", "c1\tx\tc1\t15\t2\n", "c1\t5\tc1\t15\tabc\n", "c1\t5\tc1\t15\t2\t9\n"):
        with gzip.open(p, "wt") as f:
            f.write(bad)
        with pytest.raises(ValueError):
            fio.read_contacts(str(p))
    with pytest.raises(ValueError):
        fio.read_contacts(str(tmp_path / "missing.gz"))
    bp = tmp_path / "b.gz"
    with gzip.open(bp, "wt") as f:
        f.write("c1\t5\t0.9\nc1\t5\t1.7\nc1\t15\tnan\nc1\t25\t2.5\nc1\t35\t0.4\n")
    b, _ = fio.read_biases(str(bp), ["c1"], 10, 0.5, 2.0)
    assert list(b.values) == [0.9, -1.0, -1.0, -1.0]
    assert list(fio.lookup_biases(b, np.array([0, 0, 1]), np.array([5, 45, 5]), 10)) == [0.9, -1.0, -1.0]


def test_cli_argument_handling(tmp_path, capsys):
    from fithic_b200 import fithic as cli
    a = cli.parse_args(["-i", "x.gz", "-f", "y.gz", "-o", str(tmp_path), "-r", "5000", "-p", "0", "-b", "0", "-L", "0",
                        "-x", "All", "-tL", "0.4"])
    for name in ("x.gz", "y.gz"):  # the files are probed for gzip content like the reference does
        with gzip.open(tmp_path / name, "wt") as f:
            f.write("chr1\t0\t5000\t1\t1\n")
    a.intersfile, a.fragsfile = str(tmp_path / "x.gz"), str(tmp_path / "y.gz")
    st, lib_name = cli.settings_from_args(a)
    # the reference's falsy-zero idiom: 0 means default (fithic/fithic.py:194-220)
    assert (st.noOfPasses, st.noOfBins, st.distLowThres, st.allReg, st.biasLowerBound, lib_name) == \
        (1, 100, 0, True, 0.4, "FitHiC")
    a.contactType = "bogus"
    with pytest.raises(SystemExit) as e:
        cli.settings_from_args(a)
    assert e.value.code == 2
    a.contactType = None
    a.resolution = 0  # restriction-fragment mode
    st, _ = cli.settings_from_args(a)
    assert st.resolution == 0
    a.resolution = -5
    with pytest.raises(SystemExit) as e:
        cli.settings_from_args(a)
    assert e.value.code == 2


def test_synthetic_generator_is_deterministic():
    a = synth.make_intra(2000, 100000, 9, chroms=["chr22"], with_bias=True)[0]
    b = synth.make_intra(2000, 100000, 9, chroms=["chr22"], with_bias=True)[0]
    assert np.array_equal(a.mid1, b.mid1) and np.array_equal(a.cnt, b.cnt)
    shards = synth.lpt_shards([int(s) for s in synth.genome(None)[1]], 8)
    loads = [sum(int(synth.genome(None)[1][i]) for i in s) for s in shards]
    assert sorted(i for s in shards for i in s) == list(range(24))
    assert max(loads) / (sum(loads) / 8) < 1.05


def test_host_antitonic_matches_sklearn(lib):
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(12)
    for m, noise in ((1, 0.0), (7, 1.0), (500, 0.3), (20000, 0.05)):
        y = 1e-3 * (np.arange(m) + 1.0) ** -1.1 * np.exp(noise * rng.normal(size=m))
        if m > 100:
            y[-m // 3:] = y[-m // 3:][::-1]  # a rising tail: one long cascade
        want = IsotonicRegression(increasing=False).fit_transform(np.arange(m), y)
        got = y.copy()
        _capi.check(lib.fhc_host_antitonic(_capi.dptr(got), m))
        assert np.all(np.diff(got) <= 0)
        assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_native_writer_matches_python_formatting(tmp_path):
    """fhc_io_write_significances against the reference's own formatting expression, row filters included."""
    contacts, frags, biases, raw = synth.make_intra(70_000, 100000, 8, chroms=["chr20", "chr21", "chr22"], with_bias=True,
                                                    inter_fraction=0.3)
    n = len(contacts)
    rng = np.random.default_rng(1)
    p = rng.random(n) ** 8
    p[rng.integers(0, n, 50)] = np.nan
    p[rng.integers(0, n, 50)] = 1.0
    p[rng.integers(0, n, 50)] = 1e-300
    p[rng.integers(0, n, 50)] = 0.0
    q = np.minimum(p * 3.7, 1.0)
    e = rng.random(n) * 1e4
    e[rng.integers(0, n, 100)] = 0.0
    c1, c2 = contacts.chrs & 0xffff, contacts.chrs >> 16
    b1 = fio.lookup_biases(biases, c1, contacts.mid1, 100000)
    b2 = fio.lookup_biases(biases, c2, contacts.mid2, 100000)
    for kw in (dict(), dict(allReg=True), dict(interOnly=True), dict(distLowThres=200000, distUpThres=5000000)):
        st = Settings(resolution=100000, **kw)
        a, b = str(tmp_path / "py.gz"), str(tmp_path / "native.gz")
        rows_py = fio.write_significances(a, contacts, p, q, e, b1, b2, st)
        rows_nat = fio.write_significances_native(b, contacts, p, q, e, biases, st, nthreads=3)
        assert rows_py == rows_nat
        with gzip.open(a, "rb") as fa, gzip.open(b, "rb") as fb:
            assert fa.read() == fb.read()
    # no bias table: both bias columns are 1
    st = Settings(resolution=100000)
    a, b = str(tmp_path / "py2.gz"), str(tmp_path / "native2.gz")
    ones = np.ones(n)
    fio.write_significances(a, contacts, p, q, e, ones, ones, st)
    fio.write_significances_native(b, contacts, p, q, e, None, st, nthreads=1)
    with gzip.open(a, "rb") as fa, gzip.open(b, "rb") as fb:
        assert fa.read() == fb.read()


def test_fast_double_formatting_is_printf(lib):
    """The writer's own "%e" / "%f" (one extended-precision multiply, printf only near ties) against Python's."""
    import ctypes
    import struct
    buf = ctypes.create_string_buffer(512)
    rng = np.random.default_rng(2)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 2.0 ** -11, 2.0 ** -20, 9.9999995, 9.99999949999, 9.9999995000001, 1e-7, 123456.5,
            0.1, 1e-300, 5e-324, 1e300, 1.7976931348623157e308, 999999.95, 0.0000005, 0.0000015, 1e15, 1e11 + 0.5, -3.25,
            float("nan"), float("inf"), -float("inf"), 312.808149, 4.8828125e-04, 1.2345675, 7.4e-06, 2.15549759662018e-06]
    vals += list(rng.random(20000) ** 12) + list(rng.random(5000) * 1e5) + list(10.0 ** rng.uniform(-280, 280, 20000))
    vals += [struct.unpack("<d", struct.pack("<Q", int(b)))[0] for b in rng.integers(1, 0x7fefffffffffffff, 20000)]
    vals += [k / 2.0 ** j for k in range(1, 400, 7) for j in range(0, 30, 3)]  # short dyadic values: exact ties
    for v in vals:
        v = float(v)
        n = lib.fhc_io_format_double(v, ord("e"), buf)
        assert buf.value[:n].decode() == "%e" % v, (v, buf.value, "%e" % v)
        n = lib.fhc_io_format_double(v, ord("f"), buf)
        assert buf.value[:n].decode() == "%f" % v, (v, buf.value, "%f" % v)


@pytest.mark.parametrize("case", ["null", "signal", "T_small", "ties", "offset"])
def test_bh_cut_from_value_histogram_is_exact(lib, case):
    """fhc_host_bh_cut_find (the same code the device kernel runs): every p-value at or above the cut has q == 1.0 in
    the reference's correction (fithic/myStats.py:24-48), so leaving it unranked changes nothing."""
    rng = np.random.default_rng(7)
    n = 200_000
    p = rng.random(n)
    T, off = 5.0e6, 0.0
    if case == "signal":
        p[:20_000] = rng.random(20_000) ** 6 * 1e-4
    elif case == "T_small":
        T = 1000.0  # far fewer tests than lines: the rank bound is > 1 and nothing closes early
    elif case == "ties":
        p = np.round(p, 2)
        p[p == 0] = 1e-12
    elif case == "offset":
        off = 3.0e6
    p[rng.integers(0, n, 5000)] = 1.0
    p[rng.integers(0, n, 100)] = np.nan
    p_cut0 = float(lib.fhc_bh_p_cut(T, off + n))
    hist = cut_hist_numpy(p, p_cut0)
    for v in (0.0, 1e-300, 3e-9, 0.0157, 0.5, 0.99999):
        assert lib.fhc_host_bh_cut_bucket(v) == (int(np.float64(v).view(np.uint64)) >> 47 if v > 0 else 0)
    cut = float(lib.fhc_host_bh_cut_find(_capi.dptr(hist), T, off, p_cut0))
    assert cut <= p_cut0
    # the reference's correction with ranks shifted by `off` (what a GPU holding a higher key range sees)
    order = np.argsort(p, kind="stable")
    ps = p[order]
    with np.errstate(invalid="ignore"):
        bh = np.where(ps == 1.0, 1.0, np.minimum(ps * T / (off + np.arange(1, n + 1)), 1.0))
    q = np.empty(n)
    run = 0.0
    for i in range(n):  # max(bh, prev) with Python's NaN semantics (KAT-BH4)
        run = max(bh[i], run)
        q[order[i]] = run
    with np.errstate(invalid="ignore"):
        above = p >= cut
    assert np.all(q[above & ~np.isnan(p)] == 1.0)
    if case == "null":
        assert np.sum(p < cut) < 0.01 * n  # nearly nothing is left to sort on null data
    if case == "offset":
        assert 0.5 < cut < p_cut0  # ranks start at 3e6 of T = 5e6: p T / rank reaches 1 near p = 0.6
    if case == "T_small":
        assert cut == p_cut0 or np.all(q[p >= cut] == 1.0)


def test_one_minus_exp_is_minus_expm1(lib):
    """The single-path 1 - exp(y) of the work-list kernels (cephes_dev.cuh) against libm's -expm1, including where
    the result saturates to exactly 1.0 and the sign of zero."""
    rng = np.random.default_rng(1)
    ys = -np.exp(rng.uniform(np.log(1e-300), np.log(120), 50_000))
    ys = np.concatenate([ys, [-0.0, -1e-320, -0.3465, -0.34657359, -0.346574, -1.0, -36.7, -37.0, -37.42, -37.43,
                              -37.44, -38, -745, -1e9]])
    got = np.array([lib.fhc_host_one_minus_exp(float(y)) for y in ys])
    want = -np.expm1(ys)
    assert np.array_equal(got == 1.0, want == 1.0)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-320)) < 4e-16
    assert lib.fhc_host_one_minus_exp(-0.0) == 0.0 and not np.signbit(lib.fhc_host_one_minus_exp(-0.0))


@pytest.mark.parametrize("N", [33, 34, 100, 5000, 4219169, 300_000_000, (1 << 31) - 1])
def test_short_tail_sums_in_place_equal_the_queue_bit_for_bit(lib, N):
    """Tail sums of up to 32 terms are summed inside pval_finish_kernel (tail_short_sum) instead of going through the
    queue of pval_iterate_kernel (tail_fwd_load / tail_fwd_step); the in-place form drops the renormalisation, which
    cannot trigger for that few factors: numerator and denominator must come out with the same bits, for every count the
    in-place form takes (2 ... 33) and priors on both sides of the expectation, tiny and close to 1."""
    import ctypes
    rng = np.random.default_rng(N % 7919)
    num = [ctypes.c_double(), ctypes.c_double()]
    den = [ctypes.c_double(), ctypes.c_double()]
    checked = 0
    for c in range(2, 34):
        if c - 1 >= N:  # the tail class needs count - 1 < N
            continue
        lam = np.exp(rng.uniform(np.log(0.01), np.log(100.0), 300))  # expectation relative to the count
        priors = np.minimum(c * lam / N, 1.0 - 1e-12)
        priors = np.concatenate([priors, [1e-300, 1e-15, 0.5, 1.0 - 2.0 ** -53]])
        for x in priors:
            for k in (0, 1):
                lib.fhc_host_tail_sum(c, N, float(x), 1 - k, ctypes.byref(num[k]), ctypes.byref(den[k]))
            a = np.array([num[0].value, den[0].value]).view(np.uint64)
            b = np.array([num[1].value, den[1].value]).view(np.uint64)
            assert np.array_equal(a, b), (c, N, x, num[0].value, num[1].value, den[0].value, den[1].value)
            checked += 1
    assert checked > 0


@pytest.mark.parametrize("N", [100, 171, 5000, 4219169, 300_000_000, 900_000_000, (1 << 31) - 1])
def test_list_pipeline_arithmetic_matches_oracle(lib, N):
    """fhc_host_bdtrc_lists = the source pval_front / pval_iterate / pval_finish run for one contact (series log1p +
    single-path expm1 for count 1, division-free classification, continued fraction / tail sum, prefactor with the
    divisions folded into the exponent), compiled for the host, against scipy.special.bdtrc as restated by the oracle."""
    rng = np.random.default_rng(N % 9973)
    n = 20_000
    cmax = min(N, 4000)
    cnt = np.minimum(np.floor(np.exp(rng.uniform(0, np.log(cmax + 1), n))).astype(np.int64), cmax)
    ratio = np.exp(rng.uniform(np.log(0.01), np.log(100), n))
    prior = np.minimum(cnt * ratio / N, 1.0)
    prior[::97] = np.exp(rng.uniform(np.log(1e-14), 0, len(prior[::97])))
    cnt[::5] = 1
    cnt[7::1000] = 0
    prior[3::1000] = np.nan
    prior[5::1000] = 0.0
    prior[9::1000] = -0.25
    want = O.bdtrc(cnt - 1.0, N, prior)
    got = np.array([lib.fhc_host_bdtrc_lists(int(c), N, float(x)) for c, x in zip(cnt, prior)])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got == 1.0, want == 1.0) and np.array_equal(got == 0.0, want == 0.0)
    ok = ~np.isnan(want) & ~((np.abs(want) < 1e-290) & (np.abs(got) < 1e-290)) & (want != 0)
    rel = np.abs(got[ok] - want[ok]) / np.abs(want[ok])
    assert rel.max() <= 1e-6  # the contract; count == 1 is at 4e-16
    k0 = ok & (cnt == 1)
    assert np.max(np.abs(got[k0] - want[k0]) / np.abs(want[k0])) < 2e-15


def test_known_deviation_where_glibc_log_is_not_correctly_rounded(lib):
    """lbeta cancels two terms of size N log N, so one ulp of log(N - c + 1) moves the p-value by ulp(lgam(N)) ~ 4e-6 ... 8e-6
    (SURVEY F8).  The table kernel evaluates log correctly rounded; glibc's log (what scipy's cephes calls) is not, for about
    one argument in 15,000 at some magnitudes.  Those entries -- and only those -- differ from scipy by that one ulp; the
    reference itself would change there with another libm.  Characterised here so that it is a known, bounded deviation."""
    import decimal
    import math
    ol = O._lib()
    ol.oracle_lbeta.restype = ctypes.c_double
    ol.oracle_lbeta.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.fhc_host_log_cr.restype = ctypes.c_double
    lib.fhc_host_log_cr.argtypes = [ctypes.c_double]
    decimal.getcontext().prec = 60
    N = 300_000_000
    differ = [c for c in range(1, 40_001) if lib.fhc_host_lbeta(float(c), float(N - c + 1)) !=
              ol.oracle_lbeta(float(c), float(N - c + 1))]
    assert len(differ) <= 6  # 3 with the glibc of this image (12660, 17560, 30040); 0 at N = 9e8, the bench workload
    for c in differ:
        x = N - c + 1
        ours, libm = lib.fhc_host_log_cr(float(x)), math.log(float(x))
        assert ours != libm  # the whole difference comes from log(b)
        exact = decimal.Decimal(x).ln()
        assert abs(decimal.Decimal(ours) - exact) < abs(decimal.Decimal(libm) - exact)  # and ours is the rounded one
        d = abs(lib.fhc_host_lbeta(float(c), float(N - c + 1)) - ol.oracle_lbeta(float(c), float(N - c + 1)))
        assert d <= 2 * math.ulp(N * math.log(N))  # one rounding step of the big terms: 3.8e-6 here


def test_host_lbeta_table_with_the_c_library_log_is_scipys(lib):
    """fhc_host_lbeta_table: the same cephes source with the C library's log -- bit for bit the oracle's (= scipy's) lbeta,
    the entries of the previous test included; elsewhere identical to the device table's arithmetic."""
    ol = O._lib()
    ol.oracle_lbeta.restype = ctypes.c_double
    ol.oracle_lbeta.argtypes = [ctypes.c_double, ctypes.c_double]
    for N, ntab in ((300_000_000, 40_001), ((1 << 31) - 1, 40_001), (900_000_000, 5000), (4219169, 3000), (1000, 1200)):
        tab = np.full(ntab, -7.0)
        assert lib.fhc_host_lbeta_table(N, _capi.dptr(tab), ntab, 4) == 0
        assert np.isnan(tab[0]) and (ntab <= N + 1 or np.isnan(tab[N + 1:]).all())
        hi = min(ntab, N + 1)
        want = np.array([ol.oracle_lbeta(float(c), float(N - c + 1)) for c in range(1, hi)])
        assert np.array_equal(tab[1:hi], want), N
        ours = np.array([lib.fhc_host_lbeta(float(c), float(N - c + 1)) for c in range(1, hi)])
        assert (ours != tab[1:hi]).sum() <= 6  # the misrounded-log entries (3 and 4 for the first two N, 0 otherwise)


def test_known_deviation_for_huge_counts_at_the_mode(lib):
    """Where the observed count sits on its expectation (prior ~ count / N), cephes evaluates the swapped continued fraction
    and stops after 300 iterations; from counts of ~2e5 on that is not enough (1e-3 off at 1e6).  K3 sums the short lower
    tail instead (cephes_dev.cuh), which stays at the true value.  Up to counts of 1e5 -- beyond any Hi-C bin pair -- the two
    agree within the 1e-6 contract; above, K3 follows Boost's binomial survival function, not cephes' truncated fraction."""
    import scipy.special as sp
    import scipy.stats as ss
    N = 900_000_000
    for c in (100, 1000, 10_000, 30_000, 100_000):
        for r in (0.5, 0.9, 0.99, 0.999, 1.0, 1.001, 1.01, 1.1, 2.0):
            x = r * c / N
            w, g = float(sp.bdtrc(c - 1, N, x)), lib.fhc_host_bdtrc_lists(c, N, x)
            assert abs(g - w) <= 1e-6 * abs(w) or (abs(w) < 1e-290 and abs(g) < 1e-290), (c, r, w, g)
    for c in (300_000, 1_000_000, 2_000_000):
        x = c / N
        cephes, ours, boost = float(sp.bdtrc(c - 1, N, x)), lib.fhc_host_bdtrc_lists(c, N, x), float(ss.binom.sf(c - 1, N, x))
        assert abs(ours - boost) <= 1e-5 * boost  # (the lbeta rounding noise both implementations share)
        assert abs(cephes - boost) > abs(ours - boost)
    assert abs(float(sp.bdtrc(10 ** 6 - 1, N, 10 ** 6 / N)) - float(ss.binom.sf(10 ** 6 - 1, N, 10 ** 6 / N))) > 1e-3 * 0.5


def test_chr_runs_round_trip():
    from fithic_b200.engine import chr_runs_of
    rng = np.random.default_rng(5)
    vals = rng.integers(0, 1 << 20, 40).astype(np.uint32)
    lens = rng.integers(1, 1000, 40)
    chrs = np.repeat(vals, lens)
    v, l = chr_runs_of(chrs)
    assert np.array_equal(np.repeat(v, l), chrs) and int(l.sum()) == len(chrs)
    assert np.all(v[1:] != v[:-1])
    v, l = chr_runs_of(np.zeros(0, dtype=np.uint32))
    assert len(v) == 0 and len(l) == 0


def test_varsize_frag_pairs_bit_exact_against_reference_fixture(lib):
    """fhc_host_frag_pairs_varsize (generate_FragPairs, restriction-fragment branch, fithic/fithic.py:691-778) on the
    bundled HindIII fragments: `[1]`, `[7]`, `[3]` of every bin and possibleIntraInRangeCount as the unmodified reference
    computed them, first and second pass (the second carries the outlier decrements)."""
    from tests.util import R0_CASES
    contacts, frags, biases, st, ref, _ = load_golden(R0_CASES[0])
    assert st.resolution == 0 and frags.mids is not None
    for k, r in enumerate(ref):
        nb = len(r["bins"])
        bins = dict(n=nb, lb=np.array([b["lb"] for b in r["bins"]], dtype=np.int64),
                    ub=np.array([b["ub"] for b in r["bins"]], dtype=np.int64),
                    sumcc=np.array([b["sumcc"] for b in r["bins"]], dtype=np.int64))
        dec = None
        if k > 0:  # outlier distances of the previous pass, binned like makeBinsFromInteractions :528-548
            dec = np.zeros(nb, dtype=np.int64)
            for d in ref[k - 1]["outliersdist"]:
                i = int(np.searchsorted(bins["ub"], d, side="left"))
                dec[min(i, nb - 1)] += 1
        fp = frag_pairs(lib, frags, st, bins, dec)
        assert fp["possibleIntraInRangeCount"] == r["possibleIntraInRangeCount"] == r["T"]
        assert [int(v) for v in bins["pairs"]] == [b["pairs"] for b in r["bins"]]
        assert [int(v) for v in bins["pairs7"]] == [b["pairs7"] for b in r["bins"]]
        assert [float(v) for v in bins["sumdist"]] == [b["sumdist"] for b in r["bins"]]
        x, y = calculate_probabilities(bins, r["N"])
        assert sorted(x) == list(r["x"])
        assert [v for _, v in sorted(zip(x, y))] == list(r["y"])


def _varsize_case(rng, nchr, nmax, nb, U):
    mids, off = [], [0]
    for _ in range(nchr):
        n = int(rng.integers(0, nmax))
        f = np.sort(rng.integers(0, 40_000_000, n)).astype(np.int64)
        mids.append(f)
        off.append(off[-1] + n)
    mids = np.ascontiguousarray(np.concatenate(mids) if mids else np.zeros(0, np.int64))
    off = np.asarray(off, dtype=np.int64)
    top = U if U >= 0 else 30_000_000
    ub = np.sort(rng.choice(np.arange(1000, top), nb, replace=False)).astype(np.int64)
    lb = np.concatenate([[0], ub[:-1] + 1]).astype(np.int64)
    return mids, off, lb, ub


@pytest.mark.parametrize("L,U", [(-1, -1), (20000, 5_000_000), (0, 200_000), (-1, 1_000_000), (3_000_000, -1)])
def test_varsize_frag_pairs_prefix_sums_against_the_walk(lib, L, U):
    """csrc/fragpairs.cu (the cells of the -r 0 possible-pair kernel, run serially on the host) against the pair-by-pair
    walk of fhc_host_frag_pairs_varsize: `[1]`, `[7]` and all totals exact; `[3]` -- the exact sum rounded once instead of
    a double accumulated pair by pair -- to 1e-12."""
    rng = np.random.default_rng(abs(L) * 7 + abs(U))
    for nchr, nmax, nb in ((3, 600, 20), (1, 2, 5), (4, 1500, 60), (2, 900, 1), (1, 1, 3)):
        mids, off, lb, ub = _varsize_case(rng, nchr, nmax, nb, U)
        res = {}
        for name in ("fhc_host_frag_pairs_varsize", "fhc_host_frag_pairs_varsize_prefix"):
            p1 = np.full(nb, -3, dtype=np.int64)   # (outlier decrements of a later pass ride in)
            p7 = np.full(nb, -3, dtype=np.int64)
            sd = np.zeros(nb, dtype=np.float64)
            tot = np.zeros(5, dtype=np.int64)
            check(getattr(lib, name)(dptr(mids), dptr(off), nchr, L, U, dptr(lb), dptr(ub), nb, dptr(p1), dptr(p7), dptr(sd),
                                     dptr(tot)))
            res[name] = (p1, p7, sd, tot)
        a, b = res["fhc_host_frag_pairs_varsize"], res["fhc_host_frag_pairs_varsize_prefix"]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
        assert np.allclose(a[2], b[2], rtol=1e-12, atol=0.0)
        from fithic_b200.engine import varsize_pairs_in_range
        assert varsize_pairs_in_range(mids, off, L, U) == int(a[3][0])


def test_varsize_frag_pairs_prefix_sums_on_the_reference_fixture(lib):
    """The same on the bundled HindIII fragments with the reference's own bins: `[1]` and `[7]` as the unmodified reference
    computed them, `[3]` within 1e-12 of its double accumulation."""
    from tests.util import R0_CASES
    contacts, frags, biases, st, ref, _ = load_golden(R0_CASES[0])
    r = ref[0]
    nb = len(r["bins"])
    order = sorted(range(len(frags.chroms)), key=lambda i: frags.chroms[i])
    order = [i for i in order if frags.n_mappable[i] > 0]
    mids = np.ascontiguousarray(np.concatenate([np.asarray(frags.mids[i], dtype=np.int64) for i in order]))
    off = np.zeros(len(order) + 1, dtype=np.int64)
    np.cumsum([len(frags.mids[i]) for i in order], out=off[1:])
    lb = np.array([b["lb"] for b in r["bins"]], dtype=np.int64)
    ub = np.array([b["ub"] for b in r["bins"]], dtype=np.int64)
    p1, p7, sd, tot = np.zeros(nb, np.int64), np.zeros(nb, np.int64), np.zeros(nb), np.zeros(5, np.int64)
    check(lib.fhc_host_frag_pairs_varsize_prefix(dptr(mids), dptr(off), len(order), st.L, st.U, dptr(lb), dptr(ub), nb,
                                                 dptr(p1), dptr(p7), dptr(sd), dptr(tot)))
    assert int(tot[0]) == r["possibleIntraInRangeCount"]
    assert [int(v) for v in p1] == [b["pairs"] for b in r["bins"]]
    assert [int(v) for v in p7] == [b["pairs7"] for b in r["bins"]]
    want = np.array([b["sumdist"] for b in r["bins"]])
    assert np.allclose(sd, want, rtol=1e-12, atol=0.0), np.max(np.abs(sd - want) / want)


def test_refapi_generate_fragpairs_restriction_fragments(lib, tmp_path):
    """refapi.generate_FragPairs / calculateProbabilities with resolution 0 (fithic/fithic.py:691-778, :851): the
    reference's return tuple and binStats rows from files, first pass of the bundled HindIII case (host stages only)."""
    from fithic_b200 import refapi as F, synth
    from tests.util import R0_CASES
    for name in R0_CASES:
        contacts, frags, biases, st, ref, extra = load_golden(name)
        cpath, fpath, bpath = synth.write_inputs(str(tmp_path), contacts, frags, 0, extra["bias_raw"], biases, prefix=name)
        assert (bpath is None) == (biases is None)
        F.reset()
        F.distLowThres, F.distUpThres, F.mappThres, F.noOfBins = st.distLowThres, st.distUpThres, 1, st.noOfBins
        F.interOnly, F.allReg, F.logfile = False, False, None
        F.set_resolution(cpath, 0)
        r = ref[0]
        binStats = {i: [(b["lb"], b["ub"]), 0, b["sumcc"], 0, 0, 0, [], 0] for i, b in enumerate(r["bins"])}
        (binStats, noOfFrags, maxd, T, possInter, interChrProb, base) = F.generate_FragPairs(
            r["observedInterAllCount"], r["observedInterAllSum"], binStats, fpath, 0)
        assert T == r["possibleIntraInRangeCount"] and noOfFrags == int(frags.n_mappable.sum())
        assert st.distLowThres < maxd <= st.distUpThres
        for i, b in enumerate(r["bins"]):
            assert (binStats[i][1], binStats[i][7], binStats[i][3]) == (b["pairs"], b["pairs7"], b["sumdist"])
        x, y, _ = F.calculateProbabilities({}, binStats, 0, str(tmp_path / (name + ".fithic_pass1")), r["N"])
        assert sorted(x) == list(r["x"])
        assert (tmp_path / (name + ".fithic_pass1.txt")).exists()
        if bpath:
            b = F.read_biases(bpath)
            assert b.sparse and np.array_equal(b.values, biases.values) and np.array_equal(b.mids, biases.mids)
    F.reset()


def test_sparse_bias_layout_for_restriction_fragments(tmp_path):
    """-r 0: io.read_biases builds the grid-free layout (per chromosome the loci in ascending mid order, FIRST occurrence of
    a repeated locus wins, fithic/fithic.py:823-829); lookup_biases and the native writer find loci by binary search."""
    from fithic_b200.engine import Contacts
    lines = [("chrB", 900, 1.5), ("chrA", 5000, 0.8), ("chrA", 120, 1.1), ("chrA", 5000, 1.9), ("chrB", 17, 3.0),
             ("chrA", 77777, float("nan")), ("chrB", 901, 0.7)]
    path = str(tmp_path / "bias.gz")
    with gzip.open(path, "wt") as f:
        for c, m, b in lines:
            f.write("%s\t%d\t%s\n" % (c, m, b))
    chroms = ["chrA", "chrB"]
    b, _ = fio.read_biases(path, chroms, 0, 0.5, 2.0)
    assert b.sparse and b.chr_off.tolist() == [0, 3, 6]
    assert b.mids.tolist() == [120, 5000, 77777, 17, 900, 901]
    assert b.values.tolist() == [1.1, 0.8, -1.0, -1.0, 1.5, 0.7]  # 5000 keeps its first value; NaN and 3.0 -> -1
    cid = np.array([0, 0, 0, 1, 1, 1, 5], dtype=np.int64)
    mids = np.array([5000, 121, 77777, 901, 17, 5000, 120], dtype=np.int32)
    assert fio.lookup_biases(b, cid, mids, 0).tolist() == [0.8, -1.0, -1.0, 0.7, -1.0, -1.0, -1.0]
    # the native writer against the Python writer on the same sparse table
    n = 6
    contacts = Contacts(np.array([120, 5000, 900, 17, 5000, 901], dtype=np.int32),
                        np.array([5000, 77777, 901, 900, 5001, 901], dtype=np.int32), np.arange(1, n + 1, dtype=np.int32),
                        np.array([0, 0, 1 | (1 << 16), 1 | (1 << 16), 0, 1 | (1 << 16)], dtype=np.uint32), chroms)
    st = Settings(resolution=0)
    p, q, e = np.linspace(0.1, 0.9, n), np.linspace(0.2, 1.0, n), np.linspace(0, 5, n)
    c1, c2 = contacts.chrs & 0xffff, contacts.chrs >> 16
    b1 = fio.lookup_biases(b, c1, contacts.mid1, 0)
    b2 = fio.lookup_biases(b, c2, contacts.mid2, 0)
    fio.write_significances(str(tmp_path / "py.gz"), contacts, p, q, e, b1, b2, st)
    fio.write_significances_native(str(tmp_path / "native.gz"), contacts, p, q, e, b, st, nthreads=2)
    assert gzip.open(tmp_path / "py.gz").read() == gzip.open(tmp_path / "native.gz").read()


def test_cli_shards_lines_at_chromosome_boundaries_and_probes_gzip_content(tmp_path):
    from fithic_b200 import fithic as cli
    runs = (np.zeros(4, dtype=np.uint32), np.array([400, 350, 150, 100]))
    assert cli.shard_lines(runs, 1000, 2) == [0, 500, 1000]          # no boundary within 5 % of the even cut
    assert cli.shard_lines(runs, 1000, 4) == [0, 250, 500, 750, 1000]
    runs = (np.zeros(4, dtype=np.uint32), np.array([260, 245, 240, 255]))
    assert cli.shard_lines(runs, 1000, 4) == [0, 260, 505, 745, 1000]  # cuts snapped to the chromosome boundaries
    assert cli.shard_lines(None, 10, 3) == [0, 3, 6, 10]
    assert cli.shard_lines(runs, 0, 2) == [0, 0, 0]
    # the reference opens the file with gzip and reads a line (fithic/fithic.py:139-141): content decides, not the name
    good, bad = tmp_path / "contacts.txt", tmp_path / "contacts.gz"
    with gzip.open(good, "wt") as f:
        f.write("chr1\t5000\tchr1\t15000\t3\n")
    bad.write_text("chr1\t5000\tchr1\t15000\t3\n")
    assert cli._is_gz(str(good)) and not cli._is_gz(str(bad)) and not cli._is_gz(str(tmp_path / "missing"))
