"""GPU parity of each C-ABI entry point against the oracle / numpy on seeded inputs, plus the reference KATs
(generated from the unmodified reference)."""
import ctypes

import numpy as np
import pytest

from fithic_b200 import _capi
from fithic_b200._capi import check, dptr
from oracle import fithic_oracle as O
from tests.util import rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

DEV = "cuda"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------------------------------
# K3 on inputs no reader would produce: the three implementations must agree on every line
# ---------------------------------------------------------------------------------------------------------------------
def _dirty_contacts(rng, n, res, nchr, per_chr_loci):
    """Contacts as runs of chromosome pairs with everything a front kernel has a branch for: mid points off the grid,
    negative, beyond the chromosome; chromosome ids beyond the bias table; inter lines; counts of 0, negative, 1 and huge;
    distances beyond the table."""
    run_len = rng.integers(1, 6000, size=64)
    run_len = (run_len * (n / run_len.sum())).astype(np.int64)
    run_len[-1] += n - run_len.sum()
    run_len = run_len[run_len > 0]
    c1 = rng.integers(0, nchr + 2, size=len(run_len))  # nchr, nchr + 1: not in the bias table
    c2 = np.where(rng.random(len(run_len)) < 0.7, c1, rng.integers(0, nchr + 2, size=len(run_len)))
    run_val = (c1 | (c2 << 16)).astype(np.uint32)
    chrs = np.repeat(run_val, run_len)
    span = per_chr_loci * res
    mid1 = (rng.integers(0, per_chr_loci, size=n) * res + res // 2).astype(np.int64)
    mid2 = mid1 + rng.integers(0, 400, size=n) * res
    dirty = rng.random(n)
    mid2 = np.where(dirty < 0.03, mid2 + 7, mid2)                      # off the grid
    mid1 = np.where((dirty > 0.03) & (dirty < 0.05), -mid1 - 1, mid1)  # negative
    mid2 = np.where((dirty > 0.05) & (dirty < 0.07), mid2 + span, mid2)  # beyond the chromosome's slots
    mid1 = np.where((dirty > 0.07) & (dirty < 0.08), 2_147_483_000, mid1)  # with a negative partner: distance >= 2^31
    mid2 = np.where((dirty > 0.07) & (dirty < 0.08), -2500, mid2)
    cnt = rng.geometric(0.45, size=n).astype(np.int64)
    cnt = np.where(dirty > 0.98, rng.integers(-3, 1, size=n), cnt)
    cnt = np.where((dirty > 0.96) & (dirty < 0.98), rng.integers(1000, 200_000, size=n), cnt)
    swap = rng.random(n) < 0.5
    m1 = np.where(swap, mid2, mid1)
    m2 = np.where(swap, mid1, mid2)
    return m1.astype(np.int32), m2.astype(np.int32), cnt.astype(np.int32), chrs, run_val, run_len


@pytest.mark.parametrize("mode,with_bias,regular,LU", [
    (_capi.MODE_INTRA_ONLY, True, True, (0, -1)),
    (_capi.MODE_INTRA_ONLY, True, False, (20000, 1_000_000)),
    (_capi.MODE_INTRA_ONLY, False, True, (0, -1)),
    (_capi.MODE_ALL, True, True, (0, -1)),
    (_capi.MODE_INTER_ONLY, True, True, (0, -1)),
])
def test_pvalues_on_dirty_lines_all_implementations_agree(lib, monkeypatch, mode, with_bias, regular, LU):
    """fhc_pvalues through the tile-phased kernel (the most literal restatement of fit_Spline's branch order,
    pvalue_common.cuh), the first and the second front kernel of the work-list pipeline, with the chromosome ids as an
    array and as runs, the second one also behind fhc_pvalues_prepass: p and ExpCC agree on every line -- bit for bit between
    the front kernels (same arithmetic), within 1e-12 and with the same NaN pattern against the tile kernel."""
    rng = np.random.default_rng(77)
    n, res, nchr, loci = 200_003, 5000, 5, 3000
    m1, m2, cnt, chrs, run_val, run_len = _dirty_contacts(rng, n, res, nchr, loci)
    D = 350  # shorter than the longest distance: lines beyond the table get a NaN prior
    lut = np.exp(-np.arange(D) / 60.0) * 2e-8
    lut[0] = 0.0  # a prior of exactly 0
    lut[1] = 1.0  # ... and of exactly 1 (times a bias of 1: the incbet exits)
    bias = np.exp(rng.normal(0.0, 0.4, size=nchr * loci))
    bias[rng.random(len(bias)) < 0.03] = -1.0  # discarded loci
    bias[rng.random(len(bias)) < 0.01] = 1.0
    bias[5] = np.nan
    chr_off = (np.arange(nchr + 1) * loci).astype(np.int64)
    bias_mid = (np.tile(np.arange(loci), nchr) * res + res // 2).astype(np.int32)
    bias_mid[::97] += 1  # slots read for another mid point: "missing"
    N_intra, N_inter = 40_000_000, 7_000_000
    starts = np.concatenate([[0], np.cumsum(run_len)]).astype(np.int64)
    d = {k: dev(v) for k, v in dict(m1=m1, m2=m2, cnt=cnt, chrs=chrs.view(np.int32), lut=lut, bias=bias, off=chr_off,
                                    bmid=bias_mid, rs=starts, rv=run_val.view(np.int32)).items()}
    ntab = 4096
    tabs = []
    for N in (N_intra, N_inter):
        h = np.empty(ntab)
        check(lib.fhc_host_lbeta_table(N, dptr(h), ntab, 2))
        tabs.append(dev(h))
    wsb = int(lib.fhc_pvalues_workspace_bytes(n, ntab))
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    L, U = LU
    no_lut = mode == _capi.MODE_INTER_ONLY

    def run(impl, front, runs, prepass=False):
        monkeypatch.setenv("FHC_PVAL_IMPL", impl)
        if front:
            monkeypatch.setenv("FHC_PVAL_FRONT", front)
        else:
            monkeypatch.delenv("FHC_PVAL_FRONT", raising=False)
        p = torch.full((n,), -7.0, dtype=torch.float64, device=DEV)
        e = torch.full((n,), -7.0, dtype=torch.float64, device=DEV)
        b = dptr(d["bias"]) if with_bias else None
        bm = dptr(d["bmid"]) if (with_bias and not regular) else None
        off = dptr(d["off"]) if with_bias else None
        code = b12 = None
        if prepass:
            code = torch.empty(n, dtype=torch.int32, device=DEV)
            b12 = torch.empty(n, dtype=torch.float64, device=DEV)
            check(lib.fhc_pvalues_prepass(mode, dptr(d["m1"]), dptr(d["m2"]), None if runs else dptr(d["chrs"]),
                                          dptr(d["rs"]) if runs else None, dptr(d["rv"]) if runs else None,
                                          len(run_len) if runs else 0, n, b, bm, off, nchr if with_bias else 0, 0, res, L, U,
                                          0.5, 2.0, 0, dptr(code), dptr(b12), stream()))
        check(lib.fhc_pvalues(mode, dptr(d["m1"]), dptr(d["m2"]), dptr(d["cnt"]), None if runs else dptr(d["chrs"]),
                              dptr(d["rs"]) if runs else None, dptr(d["rv"]) if runs else None, len(run_len) if runs else 0, n,
                              b, bm, off, nchr if with_bias else 0, 0, res, L, U, None if no_lut else dptr(d["lut"]),
                              0 if no_lut else D, N_intra, N_inter, 3e-9, 0.5, 2.0, dptr(tabs[0]), ntab, dptr(tabs[1]), ntab,
                              None, 0, 0.0, None, dptr(p), dptr(e), dptr(code), dptr(b12), dptr(ws), wsb, stream()))
        torch.cuda.synchronize()
        return p.cpu().numpy(), e.cpu().numpy()

    p_tile, e_tile = run("tile", None, False)
    assert not (p_tile == -7.0).any() and not (e_tile == -7.0).any()
    assert (p_tile == 1.0).any() and ((p_tile > 0) & (p_tile < 1)).any()
    if mode != _capi.MODE_INTER_ONLY and LU == (0, -1):  # every class of line is present
        assert np.isnan(p_tile).any() and (p_tile == 0.0).any()
    ref = None
    for front in ("v1", "v2"):
        for runs in (False, True):
            for prepass in ((False, True) if front == "v2" else (False,)):
                p, e = run("lists", front, runs, prepass)
                what = (front, "runs" if runs else "array", "prepass" if prepass else "")
                assert np.array_equal(np.isnan(p), np.isnan(p_tile)), what
                ok = ~np.isnan(p)
                assert np.array_equal(e, e_tile, equal_nan=True), what
                assert rel_err(p[ok], p_tile[ok]) <= 1e-12, what
                if ref is None:
                    ref = p
                assert np.array_equal(p, ref, equal_nan=True), what


# ---------------------------------------------------------------------------------------------------------------------
# K3 arithmetic: scipy.special.bdtrc
# ---------------------------------------------------------------------------------------------------------------------
def gpu_bdtrc(lib, km1, N, prior, with_table=True):
    km1 = np.ascontiguousarray(km1, dtype=np.int32)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    n = len(km1)
    out = torch.empty(n, dtype=torch.float64, device=DEV)
    tab, ntab = None, 0
    if with_table:
        ntab = int(min(max(int(km1.max()) + 2, 2), N + 1, 1 << 20))
        tab = torch.empty(ntab, dtype=torch.float64, device=DEV)
        check(lib.fhc_lbeta_table(int(N), dptr(tab), ntab, stream()))
    dk, dp = dev(km1), dev(prior)  # keep the device buffers alive across the asynchronous call
    check(lib.fhc_bdtrc(dptr(dk), int(N), dptr(dp), n, dptr(tab), ntab, dptr(out), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("N", [100, 171, 5000, 4219169, 300_000_000, 900_000_000, (1 << 31) - 1])
def test_bdtrc_matches_oracle(lib, N):
    rng = np.random.default_rng(N % 9973)
    n = 200_000
    cmax = min(N, 4000)
    cnt = np.minimum(np.floor(np.exp(rng.uniform(0, np.log(cmax + 1), n))).astype(np.int64), cmax)
    ratio = np.exp(rng.uniform(np.log(0.01), np.log(100), n))  # expected / observed
    prior = np.minimum(cnt * ratio / N, 1.0)
    prior[::97] = np.exp(rng.uniform(np.log(1e-14), 0, len(prior[::97])))
    want = O.bdtrc(cnt - 1.0, N, prior)
    got = gpu_bdtrc(lib, cnt - 1, N, prior)
    assert np.array_equal(got == 1.0, want == 1.0)
    e = rel_err(got, want)
    print("N=%d max rel err %.3e" % (N, e))
    assert e <= 1e-6  # the contract; typical 1e-10 .. 1e-7 (cephes' own error vs Boost is 1e-8 .. 1e-7)
    # the per-contact fallback (no table) must agree with the table path bit for bit
    sub = slice(0, 5000)
    assert np.array_equal(gpu_bdtrc(lib, cnt[sub] - 1, N, prior[sub], with_table=False), got[sub], equal_nan=True)


def test_bdtrc_known_answers(lib):
    """values of scipy.special.bdtrc as the reference calls it."""
    kat = [(0, 1000, 1e-3, 0.6323045752290359), (0, 1000, 0.5, 1.0), (-1, 10, 0.3, 1.0), (10, 10, 0.5, 0.0),
           (4, 4219169, 2.4e-06, 0.973042930439372), (49, 4219169, 7.4e-06, 0.0011871506001787358),
           (399, 4219169, 7.4e-05, 1.0582078639719162e-06), (2, 10 ** 9, 1e-09, 0.08030139697942389),
           (1999, 10 ** 9, 1.5e-06, 6.60069186605028e-35), (5, 100, 0.0, 0.0), (5, 100, 1.0, 1.0)]
    for k, N, p, want in kat:
        got = gpu_bdtrc(lib, [k], N, [p])[0]
        assert got == want or abs(got - want) <= 1e-6 * abs(want), (k, N, p, got, want)
    assert np.isnan(gpu_bdtrc(lib, [5], 100, [-0.1])[0])
    assert np.isnan(gpu_bdtrc(lib, [5], 100, [1.5])[0])
    assert np.isnan(gpu_bdtrc(lib, [5], 100, [np.nan])[0])
    assert np.isnan(gpu_bdtrc(lib, [11], 10, [0.5])[0])  # n < k
    # n >= 2^31 is where the reference itself breaks (int32 wrap): refused loudly
    with pytest.raises(_capi.FithicB200Error):
        gpu_bdtrc(lib, [2], 1 << 31, [1e-9])


def test_lbeta_table_bit_exact(lib):
    """Device table == host build of the same source == oracle's cephes lbeta (for N + 1 > MAXGAM)."""
    ol = O._lib()
    ol.oracle_lbeta.restype = ctypes.c_double
    ol.oracle_lbeta.argtypes = [ctypes.c_double, ctypes.c_double]
    for N in (172, 4219169, 900_000_000, (1 << 31) - 1):
        ntab = min(N, 3000) + 1
        tab = torch.empty(ntab, dtype=torch.float64, device=DEV)
        check(lib.fhc_lbeta_table(N, dptr(tab), ntab, stream()))
        torch.cuda.synchronize()
        t = tab.cpu().numpy()
        want = np.array([ol.oracle_lbeta(float(c), float(N - c + 1)) for c in range(1, ntab)])
        host = np.array([lib.fhc_host_lbeta(float(c), float(N - c + 1)) for c in range(1, ntab)])
        assert np.array_equal(t[1:], host)
        assert np.array_equal(t[1:], want), N


# ---------------------------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 3, 4097, 1_000_003])
@pytest.mark.parametrize("LU", [(0, -1), (20000, 2_000_000), (-1, 50000)])
@pytest.mark.parametrize("as_runs", [False, True], ids=["chrs", "runs"])
def test_hist_distance(lib, n, LU, as_runs):
    rng = np.random.default_rng(n + 17)
    res = 10000
    nb = 5000
    i = rng.integers(0, nb, n)
    k = np.minimum((rng.pareto(0.7, n) * 3).astype(np.int64), nb - 1 - i)
    m1 = (i * res + res // 2).astype(np.int32)
    m2 = ((i + k) * res + res // 2).astype(np.int32)
    swap = rng.random(n) < 0.5
    m1, m2 = np.where(swap, m2, m1), np.where(swap, m1, m2)
    cnt = (1 + rng.poisson(3, n)).astype(np.int32)
    if n > 10:
        cnt[rng.integers(0, n, 5)] = 0           # zero counts keep the distance "seen"
        cnt[rng.integers(0, n, 5)] = 100_000     # large counts take the global-atomic path
    c1 = rng.integers(0, 3, n).astype(np.uint32)
    c2 = np.where(rng.random(n) < 0.8, c1, rng.integers(0, 3, n)).astype(np.uint32)
    run_start = run_val = None
    nruns = 0
    if as_runs and n:
        # chromosome ids as runs (a contact file grouped by chromosome): a few long runs, some of length 1 ... 3 so that
        # groups of four lines straddle boundaries, one run boundary inside the scalar tail
        cuts = np.unique(np.concatenate([[0], rng.integers(0, n, 9), [min(n - 1, 5), min(n - 1, 6), min(n - 1, 8), n - 1]]))
        vals = np.array([(a | (b << 16)) for a, b in zip(rng.integers(0, 3, len(cuts)), rng.integers(0, 3, len(cuts)))],
                        dtype=np.uint32)
        vals[::2] = (vals[::2] & 0xffff) * 0x10001  # every other run intra
        lens = np.diff(np.concatenate([cuts, [n]]))
        chrs_full = np.repeat(vals, lens).astype(np.uint32)
        c1, c2 = chrs_full & 0xffff, chrs_full >> 16
        run_start = dev(np.concatenate([cuts, [n]]).astype(np.int64))
        run_val = dev(vals.view(np.int32))
        nruns = len(vals)
    chrs = (c1 | (c2 << 16)).astype(np.uint32)
    skip = (rng.random(n) < 0.1).astype(np.uint8) * rng.integers(1, 3, n).astype(np.uint8)
    skip_limit = n // 2
    L, U = LU
    D = nb + 1
    slots, my = (3, 1) if (n % 2) else (0, 0)  # multi-GPU layout: the largest count goes to this rank's slot
    hist = torch.empty(D, dtype=torch.int64, device=DEV)
    present = torch.empty((D + 31) // 32, dtype=torch.int32, device=DEV)
    scal = torch.full((_capi.N_SCALARS + slots,), -1, dtype=torch.int64, device=DEV)
    pad = lambda a: dev(a) if n else torch.empty(4, dtype=torch.from_numpy(a).dtype, device=DEV)
    bufs = [pad(m1), pad(m2), pad(cnt), None if nruns else pad(chrs.view(np.int32)), pad(skip)]
    check(lib.fhc_hist_distance(dptr(bufs[0]), dptr(bufs[1]), dptr(bufs[2]), dptr(bufs[3]), dptr(run_start), dptr(run_val),
                                nruns, dptr(bufs[4]), skip_limit, n, L, U, res, dptr(hist), dptr(present), D, dptr(scal),
                                slots, my, stream()))
    torch.cuda.synchronize()
    h, pr, s = hist.cpu().numpy(), present.cpu().numpy().view(np.uint32), scal.cpu().numpy()
    if slots:
        assert s[_capi.S_MAX_COUNT] == 0 and s[_capi.N_SCALARS] == 0 and s[_capi.N_SCALARS + 2] == 0
        s[_capi.S_MAX_COUNT] = s[_capi.N_SCALARS + my]
    kept = ~((skip != 0) & (np.arange(n) <= skip_limit))
    intra = (c1 == c2) & kept
    d = np.abs(m1.astype(np.int64) - m2.astype(np.int64))
    inr = intra & (d >= max(L, 0)) & ((U == -1) | (d <= U))
    want = np.zeros(D, dtype=np.int64)
    np.add.at(want, d[inr] // res, cnt[inr].astype(np.int64))
    assert np.array_equal(h, want)
    bits = np.unpackbits(pr.view(np.uint8), bitorder="little")[:D].astype(bool)
    wbits = np.zeros(D, dtype=bool)
    wbits[d[inr & (cnt <= 0)] // res] = True
    assert np.array_equal(bits, wbits)
    inter = (c1 != c2) & kept
    assert s[_capi.S_INTRA_INRANGE_SUM] == cnt[inr].astype(np.int64).sum()
    assert s[_capi.S_INTRA_ALL_SUM] == cnt[intra].astype(np.int64).sum()
    assert s[_capi.S_INTER_ALL_SUM] == cnt[inter].astype(np.int64).sum()
    assert s[_capi.S_INTER_ALL_COUNT] == inter.sum()
    assert s[_capi.S_INTRA_INRANGE_LINES] == inr.sum()
    assert s[_capi.S_INTRA_ALL_LINES] == intra.sum()
    assert s[_capi.S_MAX_COUNT] == (cnt.max() if n else 0)
    assert s[_capi.S_OFFGRID] == 0
    assert s[_capi.S_NONPOS_LINES] == (inr & (cnt <= 0)).sum()


def test_hist_offgrid_is_reported(lib):
    m1 = np.array([5000, 5000, 5000, 5000], dtype=np.int32)
    m2 = np.array([15000, 15001, 25000, 5000], dtype=np.int32)
    cnt = np.ones(4, dtype=np.int32)
    chrs = np.zeros(4, dtype=np.int32)
    hist = torch.empty(8, dtype=torch.int64, device=DEV)
    present = torch.empty(1, dtype=torch.int32, device=DEV)
    scal = torch.empty(_capi.N_SCALARS, dtype=torch.int64, device=DEV)
    bufs = [dev(m1), dev(m2), dev(cnt), dev(chrs)]
    check(lib.fhc_hist_distance(dptr(bufs[0]), dptr(bufs[1]), dptr(bufs[2]), dptr(bufs[3]), None, None, 0, None, -1, 4, 0, -1,
                                10000, dptr(hist), dptr(present), 8, dptr(scal), 0, 0, stream()))
    torch.cuda.synchronize()
    assert scal.cpu().numpy()[_capi.S_OFFGRID] == 1


# ---------------------------------------------------------------------------------------------------------------------
# K4: sort and q-values
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["onesweep", "lsd"])
def sort_impl(request, monkeypatch):
    """The radix sort has two implementations (one-sweep passes with decoupled look-back / histogram + scan + scatter
    passes); the library reads FHC_SORT on every call."""
    monkeypatch.setenv("FHC_SORT", request.param)
    return request.param


@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 250_001, 3_000_000])
def test_sort_pairs(lib, n, sort_impl):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64) * 2 + rng.integers(0, 2, n).astype(np.uint64)
    if n > 100:
        keys[rng.integers(0, n, n // 3)] = keys[rng.integers(0, n, n // 3)]  # duplicates: stability matters
        keys[:50] &= np.uint64(0xff)                                          # shared high digits
    vals = np.arange(n, dtype=np.uint32)
    ki, vi = dev(keys.view(np.int64)), dev(vals.view(np.int32))
    ko, vo = torch.empty_like(ki), torch.empty_like(vi)
    wsb = int(lib.fhc_sort_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    check(lib.fhc_sort_pairs_u64(dptr(ki), dptr(vi), dptr(ko), dptr(vo), n, dptr(ws), wsb, stream()))
    torch.cuda.synchronize()
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ko.cpu().numpy().view(np.uint64), keys[order])
    assert np.array_equal(vo.cpu().numpy().view(np.uint32), order.astype(np.uint32))
    # keys that differ in two digits only (p-values of one binade): the other passes are uniform
    keys2 = (np.uint64(0x3f50000000000000) | (rng.integers(0, 1 << 16, n, dtype=np.uint64) << np.uint64(20)))
    ki, vi = dev(keys2.view(np.int64)), dev(vals.view(np.int32))  # (the sort uses its input buffers as scratch)
    check(lib.fhc_sort_pairs_u64(dptr(ki), dptr(vi), dptr(ko), dptr(vo), n, dptr(ws), wsb, stream()))
    torch.cuda.synchronize()
    order = np.argsort(keys2, kind="stable")
    assert np.array_equal(ko.cpu().numpy().view(np.uint64), keys2[order])
    assert np.array_equal(vo.cpu().numpy().view(np.uint32), order.astype(np.uint32))


def gpu_bh(lib, p, T, rank_offset=0, carry_in=0.0):
    p = np.ascontiguousarray(p, dtype=np.float64)
    n = len(p)
    pd_ = dev(p) if n else torch.empty(1, dtype=torch.float64, device=DEV)
    q = torch.full((max(n, 1),), -7.0, dtype=torch.float64, device=DEV)
    carry = torch.zeros(1, dtype=torch.float64, device=DEV)
    ns = torch.zeros(1, dtype=torch.int64, device=DEV)
    wsb = int(lib.fhc_bh_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    check(lib.fhc_bh_qvalues(dptr(pd_), n, float(T), rank_offset, carry_in, dptr(q), dptr(carry), dptr(ns), dptr(ws), wsb,
                             stream()))
    torch.cuda.synchronize()
    return q.cpu().numpy()[:n], float(carry.item()), int(ns.item())


def test_bh_known_answers(lib, sort_impl):
    """SURVEY.md Appendix B (outputs of the unmodified myStats.benjamini_hochberg_correction)."""
    nan = float("nan")
    kat = [([0.03, 0.4, 0.7, 0.01], 10, [0.15, 1, 1, 0.1]),
           ([0.03, 0.4, 0.7, 0.01], 4, [0.06, 0.5333333333333333, 0.7, 0.04]),
           ([0.02, 0.02, 1.0, 0.5, 0.02, 1.0, 0.0], 7, [0.07, 0.07, 1.0, 0.7, 0.07, 1.0, 0.0]),
           ([0.2, nan, 0.01, 1.0, nan, 0.9], 6, [0.6000000000000001, nan, 0.06, 1.0, nan, 1]),
           ([0.01, 0.011, 0.012, 0.5], 4, [0.04, 0.04, 0.04, 0.5]),   # forward running MAX, not textbook BH
           ([1e-9, 1e-3, 0.5], 1000, [1.0000000000000002e-06, 0.5, 1])]
    for p, T, want in kat:
        got, _, _ = gpu_bh(lib, p, T)
        assert np.array_equal(got, np.array(want, dtype=np.float64), equal_nan=True), (p, T, got, want)
        assert np.array_equal(got, np.array(O.benjamini_hochberg_loop(p, T), dtype=np.float64), equal_nan=True)


@pytest.mark.parametrize("n", [0, 1, 5, 4096, 100_003, 2_000_000])
def test_bh_matches_oracle_bit_exact(lib, n, sort_impl):
    rng = np.random.default_rng(n + 5)
    p = rng.random(n) ** 3
    if n > 4:
        p[rng.integers(0, n, n // 4)] = 1.0
        p[rng.integers(0, n, n // 50 + 1)] = np.nan
        p[rng.integers(0, n, n // 10)] = p[rng.integers(0, n, n // 10)]  # ties
        p[rng.integers(0, n, 2)] = 0.0
    for T in (max(n // 3, 1), 10 * n + 7):
        got, carry, ns = gpu_bh(lib, p, T)
        want = O.benjamini_hochberg(p, T)
        assert np.array_equal(got, want, equal_nan=True)
        assert ns == int(np.sum(p < 1.0))
        if ns:
            assert carry == np.nanmax(want[p < 1.0])
    # T far above the number of lines (the usual case: possible pairs >> observed lines): almost nothing is ranked
    if n:
        got, carry, ns = gpu_bh(lib, p, 1000 * n)
        assert np.array_equal(got, O.benjamini_hochberg(p, 1000 * n), equal_nan=True)


@pytest.mark.parametrize("n", [1, 7, 100_001, 3_000_000])
def test_bh_cut_hist_matches_numpy(lib, n):
    """fhc_bh_cut_hist + fhc_host_bh_cut_find: the value histogram behind the tightened cut (bh.cu)."""
    from tests.util import cut_hist_numpy
    rng = np.random.default_rng(n)
    p = rng.random(n) ** 4
    p[rng.integers(0, n, n // 5 + 1)] = 1.0
    p[rng.integers(0, n, n // 40 + 1)] = np.nan
    p[rng.integers(0, n, 2)] = 0.0
    p[rng.integers(0, n, 2)] = 1e-310  # subnormal
    for p_cut0 in (0.3, float("inf")):
        hist = torch.zeros(_capi.BH_CUT_BUCKETS, dtype=torch.int64, device=DEV)
        pd_ = dev(p)
        check(lib.fhc_bh_cut_hist(dptr(pd_), n, p_cut0, dptr(hist), stream()))
        torch.cuda.synchronize()
        want = cut_hist_numpy(p, p_cut0)
        assert np.array_equal(hist.cpu().numpy().view(np.uint64), want)
        T = 50.0 * n
        cut = float(lib.fhc_host_bh_cut_find(dptr(want), T, 0.0, p_cut0))
        q = O.benjamini_hochberg(p, T)
        with np.errstate(invalid="ignore"):
            assert np.all(q[(p >= cut) & ~np.isnan(p)] == 1.0)


@pytest.mark.parametrize("nranks", [1, 2, 8])
def test_bh_cut_from_hists(lib, nranks):
    """The multi-GPU cut kernel on gathered histograms against the host rule on their sum, shares per rank included."""
    rng = np.random.default_rng(nranks)
    B = _capi.BH_CUT_BUCKETS
    for T, scale in ((5e6, 40), (3e9, 3), (1e3, 1000), (0.0, 5)):
        h = np.zeros((nranks, B), dtype=np.uint64)
        lo, hi = int(lib.fhc_host_bh_cut_bucket(1e-12)), int(lib.fhc_host_bh_cut_bucket(0.02))
        for r in range(nranks):
            idx = rng.integers(lo, hi, 4000)
            np.add.at(h[r], idx, rng.integers(1, scale + 1, 4000).astype(np.uint64))
        p_cut0 = 0.02
        tot = np.ascontiguousarray(h.sum(axis=0), dtype=np.uint64)
        want_cut = float(lib.fhc_host_bh_cut_find(dptr(tot), float(T), 0.0, p_cut0))
        upto = int(lib.fhc_host_bh_cut_bucket(want_cut)) if want_cut < p_cut0 else B
        want_share = h[:, :upto].sum(axis=1)
        hd = dev(h.view(np.int64).reshape(-1))
        info = torch.zeros(8 + nranks, dtype=torch.int64, device=DEV)
        my = nranks - 1
        check(lib.fhc_bh_cut_from_hists(dptr(hd), nranks, my, float(T), p_cut0, dptr(info), stream()))
        got = info.cpu().numpy()
        assert float(got[:1].view(np.float64)[0]) == want_cut, (T, want_cut)
        assert got[8:].view(np.uint64).tolist() == want_share.tolist()
        assert int(got[1]) == int(want_share.sum()) and int(got[2]) == int(want_share[my]) and int(got[3]) == int(want_share.max())


@pytest.mark.parametrize("n", [0, 3, 1025, 300_001])
def test_gather_ne_one(lib, n):
    rng = np.random.default_rng(n + 1)
    q = np.ones(n)
    k = n // 7
    where = rng.choice(n, k, replace=False) if n else np.zeros(0, dtype=np.int64)
    q[where] = rng.random(k)
    q[where[: k // 5]] = np.nan
    for cap in (max(k, 1), max(k // 2, 1)):
        idx = torch.zeros(cap, dtype=torch.int32, device=DEV)
        val = torch.zeros(cap, dtype=torch.float64, device=DEV)
        cnt = torch.zeros(1, dtype=torch.int64, device=DEV)
        qd = dev(q) if n else None
        check(lib.fhc_gather_ne_one(dptr(qd), n, cap, dptr(idx), dptr(val), dptr(cnt), stream()))
        torch.cuda.synchronize()
        assert int(cnt.item()) == k
        if k <= cap and k:
            i = idx.cpu().numpy().view(np.uint32)[:k].astype(np.int64)
            v = val.cpu().numpy()[:k]
            o = np.argsort(i)
            assert np.array_equal(i[o], np.sort(where))
            assert np.array_equal(v[o], q[np.sort(where)], equal_nan=True)


def test_bh_chained_partitions(lib):
    """rank_offset / carry_in: q of a key range computed separately equals the global result (multi-GPU contract)."""
    rng = np.random.default_rng(99)
    p = rng.random(300_000) ** 2
    T = 1_000_000
    want = O.benjamini_hochberg(p, T)
    cut = np.quantile(p, 0.37)
    lo, hi = p < cut, p >= cut
    q_lo, carry, ns = gpu_bh(lib, p[lo], T)
    q_hi, _, _ = gpu_bh(lib, p[hi], T, rank_offset=ns, carry_in=carry)
    assert np.array_equal(q_lo, want[lo]) and np.array_equal(q_hi, want[hi])


# ---------------------------------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,noise", [(5, 0.0), (300, 0.0), (300, 0.3), (1024, 1.0), (5000, 0.05), (50_000, 0.2)])
def test_spline_table(lib, m, noise):
    from scipy.interpolate import UnivariateSpline
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(m)
    res = 10000
    xk = np.sort(rng.uniform(2e4, 2e8, 60))
    yk = 1e-3 * (xk / 1e4) ** -1.1 * np.exp(noise * rng.normal(size=60))   # noisy decay -> PAVA has work to do
    ius = UnivariateSpline(xk, yk, s=min(yk) ** 2)
    t, c, k = ius._eval_args
    D = 21000
    slots = np.sort(rng.choice(np.arange(D), size=min(m, D), replace=False))
    sx = slots * res
    sx = sx[(sx >= xk.min()) & (sx <= xk.max())].astype(np.int64)
    mm = len(sx)
    splineY = ius(sx)
    want = IsotonicRegression(increasing=False).fit_transform(sx, splineY)
    tc = dev(np.concatenate([t, c]))
    table = torch.empty(mm, dtype=torch.float64, device=DEV)
    lut = torch.empty(D, dtype=torch.float64, device=DEV)
    wsb = int(lib.fhc_spline_workspace_bytes(mm))
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    nt = len(t)
    dsx = dev(sx)
    check(lib.fhc_spline_table(dptr(tc[:nt]), dptr(tc[nt:]), nt, dptr(dsx), mm, float(xk.min()), float(xk.max()),
                               res, dptr(table), dptr(lut), D, dptr(ws), wsb, stream()))
    torch.cuda.synchronize()
    got = table.cpu().numpy()
    assert np.all(np.diff(got) <= 0)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-25), np.max(np.abs(got - want) / np.abs(want))
    # the staged variant (device splev, host pooling, device lookup table) gives the same table
    y = torch.empty(mm, dtype=torch.float64, device=DEV)
    check(lib.fhc_spline_eval(dptr(tc[:nt]), dptr(tc[nt:]), nt, dptr(dsx), mm, dptr(y), stream()))
    torch.cuda.synchronize()
    yh = y.cpu().numpy()
    assert np.array_equal(yh, splineY)  # FITPACK splev, bit for bit
    check(lib.fhc_host_antitonic(dptr(yh), mm))
    assert np.allclose(yh, got, rtol=1e-13, atol=1e-25)
    lut2 = torch.empty(D, dtype=torch.float64, device=DEV)
    check(lib.fhc_spline_lut(dptr(dsx), dptr(table), mm, float(xk.min()), float(xk.max()), res, dptr(lut2), D, stream()))
    torch.cuda.synchronize()
    assert np.array_equal(lut2.cpu().numpy(), lut.cpu().numpy())
    # lookup semantics of fithic/fithic.py:1066-1068
    dl = np.clip(np.arange(D, dtype=np.float64) * res, xk.min(), xk.max())
    idx = np.minimum(np.searchsorted(sx.astype(np.float64), dl, side="left"), mm - 1)
    assert np.array_equal(lut.cpu().numpy(), got[idx])


# ---------------------------------------------------------------------------------------------------------------------
# K5
# ---------------------------------------------------------------------------------------------------------------------
def test_outlier_bin_decrements(lib):
    rng = np.random.default_rng(8)
    n = 500_003
    m1 = rng.integers(0, 10_000_000, n).astype(np.int32)
    m2 = rng.integers(0, 10_000_000, n).astype(np.int32)
    outl = ((rng.random(n) < 0.02) * rng.integers(1, 4, n)).astype(np.uint8)
    ub = np.sort(rng.choice(np.arange(1, 6_000_000), 99, replace=False)).astype(np.int64)  # last bin clamps
    dec = torch.empty(len(ub), dtype=torch.int64, device=DEV)
    bufs = [dev(m1), dev(m2), dev(outl), dev(ub)]
    check(lib.fhc_outlier_bin_decrements(dptr(bufs[0]), dptr(bufs[1]), dptr(bufs[2]), n, dptr(bufs[3]), len(ub),
                                         dptr(dec), stream()))
    torch.cuda.synchronize()
    d = np.abs(m1.astype(np.int64) - m2.astype(np.int64))
    b = np.minimum(np.searchsorted(ub, d, side="left"), len(ub) - 1)
    want = np.zeros(len(ub), dtype=np.int64)
    np.add.at(want, b, outl.astype(np.int64))
    assert np.array_equal(dec.cpu().numpy(), want)


def test_outlier_bin_decrements_beyond_the_shared_memory_table(lib):
    """-b above 2048 bins (the reference has no limit): edges and counters stay in global memory."""
    rng = np.random.default_rng(9)
    n, nb = 200_001, 5000
    m1 = rng.integers(0, 10_000_000, n).astype(np.int32)
    m2 = rng.integers(0, 10_000_000, n).astype(np.int32)
    outl = ((rng.random(n) < 0.05) * rng.integers(1, 4, n)).astype(np.uint8)
    ub = np.sort(rng.choice(np.arange(1, 6_000_000), nb, replace=False)).astype(np.int64)
    dec = torch.empty(nb, dtype=torch.int64, device=DEV)
    bufs = [dev(m1), dev(m2), dev(outl), dev(ub)]
    check(lib.fhc_outlier_bin_decrements(dptr(bufs[0]), dptr(bufs[1]), dptr(bufs[2]), n, dptr(bufs[3]), nb, dptr(dec), stream()))
    torch.cuda.synchronize()
    d = np.abs(m1.astype(np.int64) - m2.astype(np.int64))
    b = np.minimum(np.searchsorted(ub, d, side="left"), nb - 1)
    want = np.zeros(nb, dtype=np.int64)
    np.add.at(want, b, outl.astype(np.int64))
    assert np.array_equal(dec.cpu().numpy(), want)


def test_mid_range_and_digest(lib):
    rng = np.random.default_rng(10)
    for n in (1, 3, 4, 1_000_003):
        m1 = rng.integers(5000, 200_000_000, n).astype(np.int32)
        m2 = rng.integers(5000, 200_000_000, n).astype(np.int32)
        out = torch.empty(2, dtype=torch.int64, device=DEV)
        bufs = [dev(m1), dev(m2)]
        check(lib.fhc_mid_range(dptr(bufs[0]), dptr(bufs[1]), n, dptr(out), stream()))
        assert out.cpu().tolist() == [int(min(m1.min(), m2.min())), int(max(m1.max(), m2.max()))]
    # digest: shards of a file add up to the digest of the whole file, whatever the split; any change of p or q shows
    n = 300_007
    p, q = rng.random(n), rng.random(n)
    p[::97] = np.nan

    def digest(lo, hi, pp=p, qq=q):
        rl = dev(np.array([0, hi - lo], dtype=np.int64))
        rg = dev(np.array([lo], dtype=np.int64))
        out = torch.empty(2, dtype=torch.int64, device=DEV)
        a, b = dev(pp[lo:hi].copy()), dev(qq[lo:hi].copy())
        check(lib.fhc_digest_lines(dptr(a), dptr(b), hi - lo, dptr(rl), dptr(rg), 1, dptr(out), stream()))
        return out.cpu().numpy().view(np.uint64)

    whole = digest(0, n)
    with np.errstate(over="ignore"):
        parts = digest(0, 100_000) + digest(100_000, 100_001) + digest(100_001, n)
    assert np.array_equal(whole, parts)
    q2 = q.copy()
    q2[12345] = np.nextafter(q2[12345], 1.0)
    assert not np.array_equal(whole, digest(0, n, p, q2))
    # two runs: lines [0, 1000) are file lines 5000 ..., the rest file lines 0 ...
    rl = dev(np.array([0, 1000, n], dtype=np.int64))
    rg = dev(np.array([5000, 0], dtype=np.int64))
    out = torch.empty(2, dtype=torch.int64, device=DEV)
    a, b = dev(p), dev(q)
    check(lib.fhc_digest_lines(dptr(a), dptr(b), n, dptr(rl), dptr(rg), 2, dptr(out), stream()))
    assert not np.array_equal(out.cpu().numpy().view(np.uint64), whole)


# ---------------------------------------------------------------------------------------------------------------------
# range partitioning for the multi-GPU correction
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nparts", [(0, 3), (5, 2), (100_003, 5), (1_000_000, 8), (4096, 1)])
def test_partition_kernels(lib, n, nparts):
    rng = np.random.default_rng(n + nparts)
    p = rng.random(n) ** 2
    if n > 10:
        p[rng.integers(0, n, n // 10)] = 1.0
        p[rng.integers(0, n, n // 50 + 1)] = np.nan
        p[rng.integers(0, n, n // 20)] = p[rng.integers(0, n, n // 20)]
    p_cut = 0.6
    with np.errstate(invalid="ignore"):
        ranked = ~((p == 1.0) | np.isnan(p) | (p >= p_cut))
    keys = np.array([lib.fhc_bh_key_of(float(v)) for v in np.quantile(p[ranked], np.arange(1, nparts) / nparts)]
                    if ranked.any() and nparts > 1 else [0] * (nparts - 1), dtype=np.uint64)
    pd_ = dev(p) if n else torch.empty(1, dtype=torch.float64, device=DEV)
    counts = torch.zeros(nparts, dtype=torch.int64, device=DEV)
    check(lib.fhc_bh_partition_count(dptr(pd_), n, dptr(keys), nparts, p_cut, dptr(counts), stream()))
    torch.cuda.synchronize()
    kp = np.array([lib.fhc_bh_key_of(float(v)) for v in p[ranked]], dtype=np.uint64)
    part = np.searchsorted(keys, kp, side="right")
    want = np.bincount(part, minlength=nparts)
    assert np.array_equal(counts.cpu().numpy(), want)
    off = np.concatenate([[0], np.cumsum(want)[:-1]]).astype(np.int64)
    cursors = dev(off)
    send = torch.full((max(n, 1),), -5.0, dtype=torch.float64, device=DEV)
    idx = torch.zeros(max(n, 1), dtype=torch.int32, device=DEV)
    q = torch.full((max(n, 1),), -7.0, dtype=torch.float64, device=DEV)
    check(lib.fhc_bh_partition_scatter(dptr(pd_), n, dptr(keys), nparts, p_cut, dptr(cursors), dptr(send), dptr(idx),
                                       dptr(q), 0, stream()))
    torch.cuda.synchronize()
    sh, ih, qh = send.cpu().numpy(), idx.cpu().numpy().view(np.uint32), q.cpu().numpy()[:n]
    tot = int(want.sum())
    assert np.array_equal(sh[:tot], p[ih[:tot]])                      # every sent value comes from the line it names
    assert np.array_equal(np.sort(ih[:tot]), np.nonzero(ranked)[0])   # each ranked line exactly once
    for r in range(nparts):
        seg = ih[off[r]:off[r] + want[r]]
        assert np.all(part[np.searchsorted(np.nonzero(ranked)[0], seg)] == r)
    assert np.all(qh[~ranked & ~np.isnan(p)] == 1.0) and np.all(np.isnan(qh[np.isnan(p)]))
    assert np.array_equal(cursors.cpu().numpy(), off + want)
    # q back: dst[idx[j]] = src[j]
    src = dev(np.arange(tot, dtype=np.float64))
    check(lib.fhc_scatter_f64(dptr(src), dptr(idx), tot, dptr(q), stream()))
    torch.cuda.synchronize()
    assert np.array_equal(q.cpu().numpy()[ih[:tot]], np.arange(tot, dtype=np.float64))


@pytest.mark.parametrize("L,U", [(-1, -1), (20000, 5_000_000), (3_000_000, -1)])
def test_varsize_frag_pairs_kernel(lib, L, U):
    """fhc_frag_pairs_varsize (csrc/fragpairs.cu, possible pairs of restriction-fragment mode from prefix sums): the kernel
    equals its own cells run serially on the host bit for bit (integers; `[3]` is one rounding of an exact 128-bit sum), on
    small inputs also the pair-by-pair walk (`[1]`, `[7]`, totals exact, `[3]` to 1e-12), and on a chromosome set the walk
    could not finish in a test (up to 1.5e5 fragments per chromosome: 1e10 pairs without -U)."""
    from tests.test_host import _varsize_case
    rng = np.random.default_rng(5 + abs(L) + abs(U))
    for nchr, nmax, nb, walk in ((3, 600, 20, True), (1, 2, 5, True), (5, 4000, 100, True), (3, 150_000, 100, False)):
        mids, off, lb, ub = _varsize_case(rng, nchr, nmax, nb, U)
        res = {}
        names = ["fhc_frag_pairs_varsize", "fhc_host_frag_pairs_varsize_prefix"] + (["fhc_host_frag_pairs_varsize"] if walk else [])
        for name in names:
            p1 = np.full(nb, -2, dtype=np.int64)
            p7 = np.full(nb, -2, dtype=np.int64)
            sd = np.zeros(nb, dtype=np.float64)
            tot = np.zeros(5, dtype=np.int64)
            args = [dptr(mids), dptr(off), nchr, L, U, dptr(lb), dptr(ub), nb, dptr(p1), dptr(p7), dptr(sd), dptr(tot)]
            if name == "fhc_frag_pairs_varsize":
                args.append(None)
            check(getattr(lib, name)(*args))
            res[name] = (p1, p7, sd, tot)
        g, h = res["fhc_frag_pairs_varsize"], res["fhc_host_frag_pairs_varsize_prefix"]
        for a, b in zip(g, h):
            assert np.array_equal(a, b)
        if walk:
            w = res["fhc_host_frag_pairs_varsize"]
            assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and np.array_equal(g[3], w[3])
            assert np.allclose(g[2], w[2], rtol=1e-12, atol=0.0)


def test_restriction_fragment_pipeline_with_gpu_possible_pairs(lib, monkeypatch):
    """-r 0 end to end with the possible pairs from the prefix-sum kernel (FHC_VARSIZE_PAIRS=gpu) against the fixture of the
    unmodified reference: counts exact, p and q inside the 1e-6 of the spec (x moves by ~1e-13 relative, see fragpairs.cu)."""
    from tests.util import R0_CASES, load_golden, rel_err
    monkeypatch.setenv("FHC_VARSIZE_PAIRS", "gpu")
    from fithic_b200.engine import Engine
    contacts, frags, biases, st, ref, _ = load_golden(R0_CASES[0])
    eng = Engine(st, frags, biases)
    eng.upload_contacts(contacts)
    outl, stats = eng.new_outlier_state()
    for passNo, r in enumerate(ref, start=1):
        got = eng.run_pass(passNo, outl, stats)
        torch.cuda.synchronize()
        assert got["N"] == r["N"] and got["T"] == r["T"]
        assert [int(v) for v in got["bins"]["pairs"]] == [b["pairs"] for b in r["bins"]]
        assert [int(v) for v in got["bins"]["pairs7"]] == [b["pairs7"] for b in r["bins"]]
        assert np.allclose(got["x"], r["x"], rtol=1e-12, atol=0.0)
        assert rel_err(got["p"].cpu().numpy(), r["p"]) <= 1e-6 and rel_err(got["q"].cpu().numpy(), r["q"]) <= 1e-6
