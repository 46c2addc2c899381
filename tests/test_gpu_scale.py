"""GPU, BASELINE-sized inputs: properties that do not need the (slow) oracle at full size.

The oracle finishes ~1 M contacts per second, so at tens of millions of contacts parity is checked through invariants:
  * the distance histogram sums to N and equals numpy's bincount of the same device-generated input (integers, exact);
  * p depends only on (distance slot, count, bias pair): contacts that share them get bit-identical p;
  * q is a non-decreasing function of p, q >= p*T/rank, q == 1 wherever p >= lines/T, NaN pattern preserved;
  * shuffling the lines permutes p and q and changes nothing else (the sort has no order dependence);
  * a random sample of lines re-computed by the oracle from the engine's own spline table agrees to 1e-6.
"""
import numpy as np
import pytest

from fithic_b200 import synth
from fithic_b200.engine import Engine, Settings
from oracle import fithic_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


# the last two are BASELINE.json's configs 3 and 4 at their full sizes (80 M pairs, 10 kb, 2 passes; 300 M pairs, 5 kb + bias)
@pytest.mark.parametrize("n_pairs,res,passes", [(20_000_000, 10000, 2), (30_000_000, 5000, 1), (80_000_000, 10000, 2),
                                                (300_000_000, 5000, 1)])
def test_full_size_invariants(lib, n_pairs, res, passes):
    dev = torch.device("cuda", 0)
    (mid1, mid2, cnt, chrs), frags, biases, per = synth.make_intra_device(n_pairs, res, 4242, dev, mean_count=3.0,
                                                                         with_bias=(res == 5000))
    st = Settings(resolution=res, noOfBins=100, noOfPasses=passes)
    eng = Engine(st, frags, biases, device=dev)
    eng.set_contacts_device(mid1, mid2, cnt, chrs)
    outl, stats = eng.new_outlier_state()
    r = None
    for passNo in range(1, passes + 1):
        r = eng.run_pass(passNo, outl, stats)
    torch.cuda.synchronize()
    p, q, e = r["p"], r["q"], r["expcc"]
    n = n_pairs
    # histogram: exact integers
    d = (mid1.long() - mid2.long()).abs()
    slot = (d // res)
    if passes == 1:
        want = torch.bincount(slot, weights=cnt.double(), minlength=eng.D).long()
        hist, _, scal = eng.hist_distance()
        torch.cuda.synchronize()
        assert torch.equal(hist, want)
        assert int(scal[0].item()) == int(cnt.long().sum().item()) == r["N"]
    # p is a function of (slot, count, bias1*bias2)
    if biases is None:
        key = slot * 4096 + cnt.long().clamp(max=4095)
        order = torch.argsort(key)
        ks, ps = key[order], p[order]
        same = ks[1:] == ks[:-1]
        assert torch.equal(ps[1:][same], ps[:-1][same])
    # q: monotone in p, bounded below by p*T/rank, exactly 1 above lines/T
    T = float(r["T"])
    ok = ~torch.isnan(p)
    order = torch.argsort(p[ok])
    ps, qs = p[ok][order], q[ok][order]
    assert bool((qs[1:] >= qs[:-1]).all())
    rank = torch.arange(1, ps.numel() + 1, device=dev, dtype=torch.float64)
    lower = torch.clamp(ps * T / rank, max=1.0)
    assert bool((qs >= lower * (1 - 1e-15)).all())
    assert bool((q[ok][p[ok] >= n / T * (1 + 1e-9)] == 1.0).all())
    assert torch.equal(torch.isnan(p), torch.isnan(q))
    assert bool(((e == 0) | (e > 0)).all())
    # permutation invariance
    perm = torch.randperm(n, device=dev)
    eng2 = Engine(st, frags, biases, device=dev)
    eng2.set_contacts_device(mid1[perm].contiguous(), mid2[perm].contiguous(), cnt[perm].contiguous(),
                             chrs[perm].contiguous())
    o2, s2 = eng2.new_outlier_state()
    r2 = None
    for passNo in range(1, passes + 1):
        r2 = eng2.run_pass(passNo, o2, s2)
    torch.cuda.synchronize()
    assert r2["N"] == r["N"] and r2["T"] == r["T"]
    assert torch.equal(r2["p"], p[perm]) and torch.equal(r2["q"], q[perm]) and torch.equal(r2["expcc"], e[perm])
    assert torch.equal(o2, outl[perm])
    # a sample of lines against the oracle's bdtrc, with the engine's own table as the prior
    idx = torch.randint(0, n, (200_000,), device=dev)
    lut = eng._ws["lut"][:eng.D]
    prior = lut[slot[idx]]
    if biases is not None:
        bv = torch.from_numpy(biases.values).to(dev)
        off = torch.from_numpy(biases.chr_off).to(dev)
        c = (chrs[idx] & 0xffff).long()
        b1 = bv[off[c] + mid1[idx].long() // res]
        b2 = bv[off[c] + mid2[idx].long() // res]
        good = (b1 > 0) & (b2 > 0)
        prior = prior * (b1 * b2)
    else:
        good = torch.ones_like(idx, dtype=torch.bool)
    want = O.bdtrc(cnt[idx].double().cpu().numpy() - 1.0, r["N"], prior.cpu().numpy())
    got = p[idx].cpu().numpy()
    g = good.cpu().numpy()
    assert np.all(got[~g] == 1.0)
    rel = np.abs(got[g] - want[g]) / np.maximum(np.abs(want[g]), 1e-290)
    assert rel.max() <= 1e-6, rel.max()


def test_inter_only_full_size(lib):
    """BASELINE.json config 5 at full size: 100 M lines, 25 kb, -x interOnly (constant prior, global BH).  Without a bias
    file p depends on the count alone -- every count is checked against the oracle's bdtrc -- and q obeys the same order
    properties as above."""
    dev = torch.device("cuda", 0)
    n, res = 100_000_000, 25000
    (mid1, mid2, cnt, chrs), frags, _ = synth.make_inter_device(n, res, 1005, dev)
    st = Settings(resolution=res, noOfBins=100, interOnly=True)
    eng = Engine(st, frags, None, device=dev)
    eng.set_contacts_device(mid1, mid2, cnt, chrs)
    outl, stats = eng.new_outlier_state()
    r = eng.run_pass(1, outl, stats)
    torch.cuda.synchronize()
    p, q, e = r["p"], r["q"], r["expcc"]
    inter = (chrs & 0xffff) != ((chrs >> 16) & 0xffff)
    n_inter = int(inter.sum().item())
    s_inter = int(cnt[inter].long().sum().item())
    assert r["observedInterAllCount"] == n_inter and r["observedInterAllSum"] == s_inter and r["T"] == n_inter
    prior = 1.0 / n_inter
    counts = torch.unique(cnt).cpu().numpy()
    want = O.bdtrc(counts.astype(np.float64) - 1.0, s_inter, np.full(len(counts), prior))
    for c, w in zip(counts, want):
        got = torch.unique(p[cnt == int(c)]).cpu().numpy()
        assert len(got) == 1 and abs(got[0] - w) <= 1e-6 * abs(w), (c, got, w)  # intra lines take the same branch (:1098)
    assert bool((e == s_inter * prior).all())
    order = torch.argsort(p)
    ps, qs = p[order], q[order]
    assert bool((qs[1:] >= qs[:-1]).all())
    rank = torch.arange(1, n + 1, device=dev, dtype=torch.float64)
    assert bool((qs >= torch.clamp(ps * float(r["T"]) / rank, max=1.0) * (1 - 1e-15)).all())
    # ties share one q: the running max over a tie run is the value of its first rank (fithic/myStats.py:35-43)
    for c in counts[:6]:
        assert torch.unique(q[cnt == int(c)]).numel() == 1


@pytest.mark.slow
def test_config3_full_size_against_oracle(lib):
    """BASELINE.json config 3 at its full size -- whole genome, 10 kb, 80 M contact pairs, two spline passes -- against the
    oracle on ALL lines: totals, observed distances, bins, possible pairs, x / y and the outlier multiset exact, the spline
    table to 1e-12, p and q to 1e-6.  A few minutes of oracle time on the box's host cores (FHC_SLOW=1); the log of the
    round's run is profiles/c3_full_oracle_rNN.log."""
    import time
    from tests.util import compare_pass, oracle_inputs
    n, res = 80_000_000, 10000
    t0 = time.time()
    contacts, frags, biases, _ = synth.make_intra(n, res, seed=1003, mean_count=4.0, with_bias=False)
    st = Settings(resolution=res, noOfBins=100, noOfPasses=2)
    eng = Engine(st, frags, biases)
    eng.upload_contacts(contacts)
    outl, stats = eng.new_outlier_state()
    got = []
    for passNo in (1, 2):
        r = eng.run_pass(passNo, outl, stats)
        torch.cuda.synchronize()
        for k in ("p", "q", "expcc"):
            r[k] = r[k].cpu().numpy().copy()
        r["outl"] = outl.cpu().numpy().copy()
        got.append(r)
    t1 = time.time()
    oc, fchr, fmid, fh, ost, ob = oracle_inputs(contacts, frags, st, biases)
    want = O.run_pipeline(oc, fchr, fmid, fh, ost, ob)
    t2 = time.time()
    assert len(want) == 2
    for r, o in zip(got, want):
        errs = compare_pass(r, o)
        lines = np.repeat(np.arange(n), r["outl"])
        assert np.array_equal(lines, np.asarray(o["outliersline"], dtype=np.int64))
        print("config 3 full size, pass %d: N %d, T %d, %d bins, %d observed distances, max rel err %s, %d outlier entries, "
              "%d lines with q < 1" % (r["passNo"], r["N"], r["T"], r["bins"]["n"], len(r["dists"]), errs, len(lines),
                                       int((r["q"] < 1).sum())))
    print("generate + engine %.1f s, oracle %.1f s" % (t1 - t0, t2 - t1))
