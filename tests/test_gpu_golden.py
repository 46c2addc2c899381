"""GPU parity against fixtures captured from the UNMODIFIED reference (tests/golden/, no oracle in the loop)."""
import os
import gzip

import numpy as np
import pytest

from fithic_b200 import synth
from tests.test_gpu_pipeline import run_engine
from tests.util import GOLDEN_CASES, GOLDEN_DIR, R0_CASES, REAL_CASES, compare_pass, load_golden, load_kat

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("pval_impl")]


@pytest.mark.parametrize("name", GOLDEN_CASES + REAL_CASES + R0_CASES)
def test_engine_matches_reference_fixture(lib, name):
    contacts, frags, biases, st, ref, _ = load_golden(name)
    got = run_engine(contacts, frags, biases, st)
    assert len(got) == len(ref)
    for r, o in zip(got, ref):
        errs = compare_pass(r, o, tol=1e-6)
        lines = np.repeat(np.arange(len(contacts)), r["outl"])
        assert np.array_equal(lines, o["outliersline"])
        print(name, "pass", r["passNo"], errs)


def test_bdtrc_reference_vectors(lib):
    from tests.test_gpu_kernels import gpu_bdtrc
    for k, N, p, want in load_kat()["bdtrc"]:
        got = gpu_bdtrc(lib, [k], N, [p])[0]
        if np.isnan(want):
            assert np.isnan(got)
        elif want in (0.0, 1.0):
            assert got == want, (k, N, p, got, want)
        else:
            assert abs(got - want) <= 1e-6 * abs(want), (k, N, p, got, want)


def test_bh_reference_vectors(lib):
    from tests.test_gpu_kernels import gpu_bh
    for p, T, want in load_kat()["bh"]:
        got, _, _ = gpu_bh(lib, p, T)
        assert np.array_equal(got, np.array(want, dtype=np.float64), equal_nan=True)


@pytest.mark.parametrize("name", ["intra_bias_LU_p2", "all_bias", "inter_only_bias"])
def test_cli_output_file_matches_reference(lib, name, tmp_path):
    """The `fithic` CLI on gz inputs: same file names, header and rows (to the 7 printed digits) as the reference."""
    from fithic_b200 import fithic as cli
    contacts, frags, biases, st, ref, extra = load_golden(name)
    cpath, fpath, bpath = synth.write_inputs(str(tmp_path), contacts, frags, st.resolution, extra["bias_raw"], biases,
                                             prefix=name)
    argv = ["-i", cpath, "-f", fpath, "-o", str(tmp_path / "out"), "-r", str(st.resolution), "-l", name,
            "-b", str(st.noOfBins), "-p", str(st.noOfPasses)]
    if st.L:
        argv += ["-L", str(st.L)]
    if st.U != -1:
        argv += ["-U", str(st.U)]
    if st.allReg:
        argv += ["-x", "All"]
    if st.interOnly:
        argv += ["-x", "interOnly"]
    if bpath:
        argv += ["-t", bpath]
    cli.main(argv)
    npass = len(ref)
    sig = tmp_path / "out" / ("%s.spline_pass%d.res%d.significances.txt.gz" % (name, npass, st.resolution))
    assert sig.exists()
    assert (tmp_path / "out" / ("%s.fithic_pass%d.res%d.txt" % (name, npass, st.resolution))).exists()
    assert (tmp_path / "out" / (name + ".fithic.log")).exists()
    # the log: the reference's own, line for line (tests/golden/make_golden_log.py)
    with open(tmp_path / "out" / (name + ".fithic.log")) as f:
        got_log = f.read().replace(str(tmp_path / "out"), "OUT")
    with open(os.path.join(GOLDEN_DIR, name + ".fithic.log")) as f:
        want_log = f.read()
    assert got_log.splitlines() == want_log.splitlines()
    with gzip.open(sig, "rt") as f:
        lines = f.readlines()
    assert len(lines) - 1 == extra["sig_nrows"]
    want = extra["sig_head"]
    assert lines[0] == want[0]
    for a, b in zip(lines[1:len(want)], want[1:]):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:5] == fb[:5], (a, b)
        for x, y in zip(fa[5:], fb[5:]):
            x, y = float(x), float(y)
            assert (np.isnan(x) and np.isnan(y)) or abs(x - y) <= 2e-6 * max(abs(y), 1e-300) + 1e-6 * (abs(y) < 1e-290), (a, b)


def test_cli_restriction_fragment_mode(lib, tmp_path):
    """`fithic -r 0` on restriction-fragment data (fithic/tests/run_tests-git.sh:28-30): file names without the `.res` part
    (fithic/fithic.py:851, :1171) and the rows of the reference's own output."""
    from fithic_b200 import fithic as cli
    name = R0_CASES[0]
    contacts, frags, biases, st, ref, extra = load_golden(name)
    cpath = str(tmp_path / "contacts.gz")
    fpath = str(tmp_path / "frags.gz")
    with gzip.open(cpath, "wt") as f:
        c1, c2 = contacts.chrs & 0xffff, contacts.chrs >> 16
        for i in range(len(contacts)):
            f.write("%s\t%d\t%s\t%d\t%d\n" % (contacts.chroms[c1[i]], contacts.mid1[i], contacts.chroms[c2[i]],
                                              contacts.mid2[i], contacts.cnt[i]))
    with gzip.open(fpath, "wt") as f:
        for ci, ch in enumerate(frags.chroms):
            for m in frags.mids[ci]:
                f.write("%s\t0\t%d\t1\t1\n" % (ch, m))
    cli.main(["-i", cpath, "-f", fpath, "-o", str(tmp_path / "out"), "-r", "0", "-l", name, "-b", str(st.noOfBins), "-p",
              str(st.noOfPasses), "-L", str(st.L), "-U", str(st.U)])
    npass = len(ref)
    sig = tmp_path / "out" / ("%s.spline_pass%d.significances.txt.gz" % (name, npass))
    assert sig.exists() and (tmp_path / "out" / ("%s.fithic_pass%d.txt" % (name, npass))).exists()
    with gzip.open(sig, "rt") as f:
        lines = f.readlines()
    assert len(lines) - 1 == extra["sig_nrows"]
    want = extra["sig_head"]
    assert lines[0] == want[0]
    for a, b in zip(lines[1:len(want)], want[1:]):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:5] == fb[:5], (a, b)
        for x, y in zip(fa[5:], fb[5:]):
            x, y = float(x), float(y)
            assert abs(x - y) <= 2e-6 * max(abs(y), 1e-300) + 1e-6 * (abs(y) < 1e-290), (a, b)
