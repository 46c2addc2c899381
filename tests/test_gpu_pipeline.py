"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded synthetic inputs."""
import numpy as np
import pytest

from fithic_b200 import synth
from fithic_b200.engine import Engine, Settings
from oracle import fithic_oracle as O
from tests.util import compare_pass, oracle_inputs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


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
