"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded synthetic inputs."""
import numpy as np
import pytest

from fithic_b200 import synth
from fithic_b200.engine import Engine, Settings
from oracle import fithic_oracle as O
from tests.util import compare_pass, oracle_inputs

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("pval_impl")]


def run_engine(contacts, frags, biases, st):
    eng = Engine(st, frags, biases)
    eng.upload_contacts(contacts)
    outl, stats = eng.new_outlier_state()
    res = []
    for passNo in range(1, st.noOfPasses + 1):
        if passNo > 1 and st.interOnly:
            break
        r = eng.run_pass(passNo, outl, stats)
        torch.cuda.synchronize()
        for k in ("p", "q", "expcc"):
            r[k] = r[k].cpu().numpy().copy()
        if "table_dev" in r:
            r["table"] = r["table_dev"].cpu().numpy().copy()
        r["outl"] = outl.cpu().numpy().copy()
        r["outl_stats"] = stats.cpu().numpy().copy()
        res.append(r)
    return res


CASES = [
    # name, n_pairs, res, chroms, mean_count, bias, inter_fraction, settings
    ("chr1_40kb", 200_000, 40000, ["chr1"], 8.0, False, 0.0, dict(noOfBins=100)),
    ("chr1_40kb_LU", 200_000, 40000, ["chr1"], 8.0, False, 0.0, dict(noOfBins=50, distLowThres=80000, distUpThres=5000000)),
    ("wg_100kb_bias_p2", 300_000, 100000, None, 4.0, True, 0.0, dict(noOfBins=100, noOfPasses=2)),
    ("wg_100kb_all", 200_000, 100000, None, 3.0, True, 0.3, dict(noOfBins=100, allReg=True)),
    ("wg_100kb_inter", 200_000, 100000, None, 3.0, False, 0.9, dict(noOfBins=100, interOnly=True)),
    ("three_chr_10kb_p3", 150_000, 10000, ["chr20", "chr21", "chr22"], 3.0, True, 0.1, dict(noOfBins=200, noOfPasses=3)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_pipeline_matches_oracle(lib, case):
    name, n, res, chroms, mean_count, bias, inter, sk = case
    contacts, frags, biases, _ = synth.make_intra(n, res, seed=1000 + len(name), chroms=chroms, mean_count=mean_count,
                                                  with_bias=bias, inter_fraction=inter)
    st = Settings(resolution=res, **sk)
    got = run_engine(contacts, frags, biases, st)
    oc, fchr, fmid, fh, ost, ob = oracle_inputs(contacts, frags, st, biases)
    want = O.run_pipeline(oc, fchr, fmid, fh, ost, ob)
    assert len(got) == len(want)
    for r, o in zip(got, want):
        errs = compare_pass(r, o)
        # outlier set: identical multiset (multiplicities accumulate over passes like the reference's SortedList)
        lines = np.repeat(np.arange(len(contacts)), r["outl"])
        assert np.array_equal(lines, np.asarray(o["outliersline"], dtype=np.int64)), name
        print(name, "pass", r["passNo"], errs, "outliers", len(lines))


def test_api_sliced_pvalues_match_single_launch(lib):
    """api.significance scores large inputs slice by slice (device->host copies overlap the next slice); results,
    outlier multiplicities and the first-duplicate statistic must not depend on the slicing."""
    from fithic_b200 import api
    from fithic_b200.engine import Contacts
    dev = torch.device("cuda", 0)
    n = (1 << 22) + 12345
    (mid1, mid2, cnt, chrs), frags, biases, _ = synth.make_intra_device(n, 10000, 77, dev, mean_count=3.0, with_bias=True)
    st = Settings(resolution=10000, noOfBins=100, noOfPasses=2)
    eng = Engine(st, frags, biases, device=dev)
    eng.set_contacts_device(mid1, mid2, cnt, chrs)
    outl, stats = eng.new_outlier_state()
    ref = None
    for passNo in (1, 2):
        ref = eng.run_pass(passNo, outl, stats)
    torch.cuda.synchronize()
    want = {k: ref[k].cpu().numpy().copy() for k in ("p", "q", "expcc")}
    want_outl, want_stats = outl.cpu().numpy().copy(), stats.cpu().numpy().copy()
    host = Contacts(mid1.cpu().numpy(), mid2.cpu().numpy(), cnt.cpu().numpy(), chrs.cpu().numpy().view(np.uint32),
                    list(frags.chroms))
    eng2 = Engine(st, frags, biases, device=dev)
    got = api.significance(host, frags, st, biases, engine=eng2)
    assert len(got) == 2
    for k in ("p", "q", "expcc"):
        assert np.array_equal(got[-1][k], want[k], equal_nan=True), k
    assert got[-1]["N"] == ref["N"] and got[-1]["T"] == ref["T"]


def test_api_q_dense_and_sparse_paths(lib):
    """api.significance brings q to the host as (line, value) pairs where q != 1.0, or whole when those are more than
    n / 128 of the lines; both routes must reproduce the device array (NaN included)."""
    from fithic_b200 import api
    # allReg with bias: inter lines with a discarded locus give p = q = NaN (fithic/fithic.py:1099-1108) -> dense route
    name, n, res, chroms, mean_count, bias, inter, sk = CASES[3]
    contacts, frags, biases, _ = synth.make_intra(n, res, seed=1000 + len(name), chroms=chroms, mean_count=mean_count,
                                                  with_bias=bias, inter_fraction=inter)
    st = Settings(resolution=res, **sk)
    want = run_engine(contacts, frags, biases, st)
    got = api.significance(contacts, frags, st, biases)
    assert got[-1]["q_exceptions"] > max(n // 128, 1024)
    for k in ("p", "q", "expcc"):
        assert np.array_equal(got[-1][k], want[-1][k], equal_nan=True), k
    assert got[-1]["q_exceptions"] == int(np.sum(want[-1]["q"] != 1.0))
    # fewer lines than the smallest pair capacity (1024): the sparse route whatever the data
    n, res = 900, 40000
    contacts, frags, biases, _ = synth.make_intra(n, res, seed=4242, chroms=["chr1"], mean_count=6.0, with_bias=True)
    st = Settings(resolution=res, noOfBins=20)
    want = run_engine(contacts, frags, biases, st)
    got = api.significance(contacts, frags, st, biases)
    n_ex = int(np.sum(want[-1]["q"] != 1.0))  # NaN != 1.0
    assert 0 < n_ex <= 1024 and got[-1]["q_exceptions"] == n_ex
    for k in ("p", "q", "expcc"):
        assert np.array_equal(got[-1][k], want[-1][k], equal_nan=True), k


def test_api_reused_host_buffers(lib):
    """HostBuffers remember which lines of q differed from 1.0; a second call with other data must reset exactly those."""
    from fithic_b200 import api
    n, res = 900, 40000
    out = api.HostBuffers(n)
    for seed in (1, 2, 3):
        contacts, frags, biases, _ = synth.make_intra(n, res, seed=seed, chroms=["chr1"], mean_count=6.0, with_bias=True)
        st = Settings(resolution=res, noOfBins=20)
        want = run_engine(contacts, frags, biases, st)
        if seed == 2:  # chromosome ids in run-length form: expanded on the device instead of uploaded
            from fithic_b200.engine import chr_runs_of
            contacts.chr_runs = chr_runs_of(contacts.chrs)
        got = api.significance(contacts, frags, st, biases, out=out)
        for k in ("p", "q", "expcc"):
            assert np.array_equal(got[-1][k], want[-1][k], equal_nan=True), (seed, k)
    out.q.fill_(0.5)  # the caller scribbles over q ...
    out.invalidate()  # ... and says so
    got = api.significance(contacts, frags, st, biases, out=out)
    assert np.array_equal(got[-1]["q"], want[-1]["q"], equal_nan=True)


@pytest.mark.parametrize("mode", ["intra_bias", "all", "inter_only"])
def test_prepass_changes_nothing(lib, monkeypatch, mode):
    """K3 behind a pre-pass (bias products, line classes and distance slots computed while the host fits) and K3 on its own
    give the same p-values, ExpCC and outlier marks bit for bit; so do chromosome ids as runs and as an array."""
    monkeypatch.setenv("FHC_PVAL_IMPL", "lists")
    kw = dict(intra_bias=dict(with_bias=True, inter_fraction=0.0), all=dict(with_bias=True, inter_fraction=0.3),
              inter_only=dict(with_bias=False, inter_fraction=0.9))[mode]
    sk = dict(intra_bias=dict(noOfPasses=2, distLowThres=200000, distUpThres=30000000), all=dict(allReg=True),
              inter_only=dict(interOnly=True))[mode]
    contacts, frags, biases, _ = synth.make_intra(250_003, 100000, seed=91, mean_count=4.0, **kw)
    order = np.argsort(contacts.chrs, kind="stable")  # grouped by chromosome pair: a few runs
    from fithic_b200.engine import Contacts, chr_runs_of
    contacts = Contacts(contacts.mid1[order], contacts.mid2[order], contacts.cnt[order], contacts.chrs[order], contacts.chroms)
    contacts.chr_runs = chr_runs_of(contacts.chrs)
    st = Settings(resolution=100000, noOfBins=100, **sk)
    got = {}
    for pre in ("1", "0"):
        for runs in ("1", "0"):
            monkeypatch.setenv("FHC_PREPASS", pre)
            monkeypatch.setenv("FHC_Q_PREFILL", pre)  # (q = 1.0 written ahead of K4 or by K4 itself)
            monkeypatch.setenv("FHC_CHR_RUNS", runs)
            got[pre, runs] = run_engine(contacts, frags, biases, st)
    monkeypatch.setenv("FHC_PREPASS", "1")
    monkeypatch.setenv("FHC_CHR_RUNS", "1")
    monkeypatch.setenv("FHC_PREPASS_LINES", "102400")  # only the first lines behind the pre-pass: K3 in two parts
    got["partial", "1"] = run_engine(contacts, frags, biases, st)
    monkeypatch.delenv("FHC_PREPASS_LINES")
    ref = got["0", "0"]
    for key, res in got.items():
        assert len(res) == len(ref)
        for a, b in zip(res, ref):
            for k in ("p", "q", "expcc", "outl"):
                assert np.array_equal(a[k], b[k], equal_nan=True), (key, k)
            assert a["N"] == b["N"] and a["T"] == b["T"]
