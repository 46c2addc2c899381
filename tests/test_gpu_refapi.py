"""GPU: the reference's function-level interface (fithic_b200/refapi.py, myStats.py) driven exactly like the reference's
main() drives its own functions (fithic/fithic.py:317-376), against fixtures captured from the unmodified reference."""
import gzip
import os

import numpy as np
import pytest
from sortedcontainers import SortedList

from fithic_b200 import synth
from tests.util import R0_CASES, load_golden, load_kat, rel_err

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("pval_impl")]


def test_benjamini_hochberg_correction_docstring_example(lib):
    from fithic_b200 import myStats
    assert myStats.benjamini_hochberg_correction([0.03, 0.4, 0.7, 0.01], 10) == [0.15, 1, 1, 0.1]
    for p, T, want in load_kat()["bh"]:
        got = myStats.benjamini_hochberg_correction(p, T)
        assert np.array_equal(np.array(got), np.array(want, dtype=np.float64), equal_nan=True)


@pytest.mark.parametrize("name", ["intra_bias_LU_p2", "all_bias", "intra_p3"] + R0_CASES)
def test_reference_main_loop_on_refapi(lib, name, tmp_path):
    from fithic_b200 import refapi as F
    contacts, frags, biases, st, ref, extra = load_golden(name)
    cpath, fpath, bpath = synth.write_inputs(str(tmp_path), contacts, frags, st.resolution, extra["bias_raw"], biases,
                                             prefix=name)
    F.reset()
    # what main() does with its globals (fithic/fithic.py:193-307)
    F.distLowThres, F.distUpThres = st.distLowThres, st.distUpThres
    F.interOnly, F.allReg, F.mappThres = st.interOnly, st.allReg, 1
    F.biasLowerBound, F.biasUpperBound = 0.5, 2
    F.noOfBins = st.noOfBins
    F.logfile = str(tmp_path / "run.log")
    res, out = st.resolution, str(tmp_path / "lib")
    tag = ".res%d" % res if res else ""  # restriction-fragment mode (resolution 0) drops the `.res` part (:851, :1171)
    F.set_resolution(cpath, res)
    outliersline, outliersdist = SortedList(), SortedList()
    for passNo in range(1, st.noOfPasses + 1):
        o = ref[passNo - 1]
        (mainDic, oic, ois, oias, N) = F.read_Interactions(cpath, bpath, outliersline if passNo > 1 else None)
        binStats = F.makeBinsFromInteractions(mainDic, st.noOfBins, N, outliersdist if passNo > 1 else None)
        (binStats, noOfFrags, maxd, T_intra, possInter, interChrProb, base) = F.generate_FragPairs(oic, ois, binStats,
                                                                                                   fpath, res)
        biasDic = F.read_biases(bpath) if bpath else 0
        if passNo == 1 and bpath:
            # biases and fragments are known now: the reference re-reads the contacts in every pass, we re-bind once
            (mainDic, oic, ois, oias, N) = F.read_Interactions(cpath, bpath, None)
        (x, y, yerr) = F.calculateProbabilities(mainDic, binStats, res, out + ".fithic_pass%d" % passNo, N)
        r = F.fit_Spline(mainDic, x, y, yerr, cpath, out + ".spline_pass%d" % passNo, biasDic, outliersline, outliersdist,
                         N, T_intra, possInter, oic, oias, ois, 0.5, 2, res, passNo)
        assert (N, oic, ois, oias) == (o["N"], o["observedInterAllCount"], o["observedInterAllSum"],
                                       o["observedIntraAllSum"])
        assert {int(k): v[1] for k, v in mainDic.items()} == dict(zip(o["dists"].tolist(), o["sums"].tolist()))
        assert len(binStats) == len(o["bins"])
        for i, b in enumerate(o["bins"]):
            assert (binStats[i][0], binStats[i][1], binStats[i][2], binStats[i][3], binStats[i][7]) == \
                ((b["lb"], b["ub"]), b["pairs"], b["sumcc"], b["sumdist"], b["pairs7"])
        assert T_intra == o["possibleIntraInRangeCount"]
        assert sorted(x) == list(o["x"])
        assert r[0] == o["splineX"].tolist()
        assert rel_err(r[1], o["newSplineY"]) <= 1e-12
        last = F.last_results(cpath)
        assert last["T"] == o["T"]
        assert rel_err(last["p"], o["p"]) <= 1e-6 and rel_err(last["q"], o["q"]) <= 1e-6
        assert list(outliersline) == o["outliersline"].tolist()
        assert list(outliersdist) == o["outliersdist"].tolist()
        assert os.path.exists(out + ".spline_pass%d%s.significances.txt.gz" % (passNo, tag))
        assert os.path.exists(out + ".fithic_pass%d%s.txt" % (passNo, tag))
    with gzip.open(out + ".spline_pass%d%s.significances.txt.gz" % (st.noOfPasses, tag), "rt") as f:
        assert len(f.readlines()) - 1 == extra["sig_nrows"]
