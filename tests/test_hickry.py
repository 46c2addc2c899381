"""KR bias computation: oracle/hickry_oracle.py against fixtures captured from the unmodified reference
fithic/utils/HiCKRy.py (CPU), and fithic_b200/hickry.py on the GPU against both."""
import gzip
import os

import numpy as np
import pytest

from oracle import hickry_oracle as K

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["hickry_pfal_10kb", "hickry_hesc_40kb", "hickry_pfal_10kb_x10"]


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_bias(name):
    """Bit for bit: removed loci, iteration counts and every bias value (the hESC case stops at the reference's cap of 30
    outer iterations without converging and is reproduced all the same)."""
    g = load(name)
    bias, info = K.return_bias(g["x"], g["y"], g["z"], int(g["n"]), float(g["perc"]))
    assert np.array_equal(info["removed"], g["removed"])
    assert (info["outer"], info["inner"]) == (int(g["outer"]), int(g["inner"]))
    assert np.array_equal(bias, g["bias"])


def test_host_side_of_hickry_matches_reference_fixture(tmp_path):
    """The host pieces of fithic_b200/hickry.py (row removal rule, bias vector, -1 insertion, output format)."""
    pytest.importorskip("torch")
    from fithic_b200 import hickry as H
    g = load("hickry_pfal_10kb")
    n = int(g["n"])
    A = K.raw_matrix(g["x"], g["y"], g["z"], n)
    removed = H.removeZeroDiagonalCSR(np.asarray(A.sum(axis=0)).reshape(-1), float(g["perc"]))
    assert np.array_equal(removed, g["removed"])
    keep = np.ones(n, dtype=bool)
    keep[removed] = False
    xs, _, _ = K.knight_ruiz(A[np.nonzero(keep)[0]][:, np.nonzero(keep)[0]].tocsr())
    bias = H.addZeroBiases(removed, H.computeBiasVector(xs))
    assert np.array_equal(bias.reshape(-1), g["bias"])
    rev = list(zip(g["chroms"].tolist(), g["mids"].tolist()))
    out = tmp_path / "bias.gz"
    H.outputBias(bias, rev, str(out))
    lines = gzip.open(out, "rt").read().splitlines()
    assert len(lines) == n
    c, m, v = lines[5].split("\t")
    assert (c, int(m)) == rev[5] and v == "%s" % np.float64(g["bias"][5])  # the reference's "%s" of a numpy float64
    with pytest.raises(RuntimeError):
        import torch
        if torch.cuda.is_available():
            raise RuntimeError("GPU present")
        H.KRDevice(g["x"], g["y"], g["z"], n)  # no CPU fallback


def write_files(g, tmp_path):
    """The fixture's loci and lines as the two gz files HiCKRy reads (fragments: chr, 0, mid, hits, mappable)."""
    chroms, mids = g["chroms"].tolist(), g["mids"].tolist()
    fpath, ipath = str(tmp_path / "frags.gz"), str(tmp_path / "contacts.gz")
    with gzip.open(fpath, "wt", compresslevel=1) as f:
        f.write("".join("%s\t0\t%d\t1\t1\n" % (c, m) for c, m in zip(chroms, mids)))
    with gzip.open(ipath, "wt", compresslevel=1) as f:
        f.write("".join("%s\t%d\t%s\t%d\t%d\n" % (chroms[a], mids[a], chroms[b], mids[b], z) for a, b, z in
                        zip(g["x"].tolist(), g["y"].tolist(), g["z"].tolist())))
    return ipath, fpath


def test_loader_numbers_loci_in_fragment_file_order(tmp_path):
    """loadfastfithicInteractions (HiCKRy.py:18-52): files -> the (x, y, z, n) the fixture was captured with; a contact on a
    locus the fragments file does not list is a KeyError like in the reference."""
    pytest.importorskip("torch")
    from fithic_b200 import hickry as H
    g = load("hickry_pfal_10kb")
    ipath, fpath = write_files(g, tmp_path)
    (x, y, z, n), rev = H.loadfastfithicInteractions(ipath, fpath)
    assert n == int(g["n"]) and rev == list(zip(g["chroms"].tolist(), g["mids"].tolist()))
    assert np.array_equal(x, g["x"]) and np.array_equal(y, g["y"]) and np.array_equal(z, g["z"])
    with gzip.open(ipath, "at") as f:
        f.write("chrNowhere\t5\tchrNowhere\t15\t1\n")
    with pytest.raises(KeyError):
        H.loadfastfithicInteractions(ipath, fpath)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_kr_spmv_matches_scipy(lib):
    torch = pytest.importorskip("torch")
    from fithic_b200 import hickry as H
    rng = np.random.default_rng(3)
    n, nnz = 5000, 400_000
    x = np.sort(rng.integers(0, n, nnz)).astype(np.int32)          # file order: grouped by row
    y = rng.integers(0, n, nnz).astype(np.int32)
    y[::7] = x[::7]                                                  # diagonal lines count twice
    z = rng.integers(1, 50, nnz).astype(np.float64)
    dev = H.KRDevice(x, y, z, n)
    A = K.raw_matrix(x, y, z, n)
    for kept in (np.arange(n), np.sort(rng.choice(n, n - 300, replace=False))):
        dev.set_kept(kept)
        v = rng.random(len(kept)) + 0.5
        out = dev.vec()
        dev.spmv(torch.from_numpy(v).to(dev.device), out)
        want = A[kept][:, kept].dot(v)
        got = out.cpu().numpy()
        assert np.max(np.abs(got - want) / np.abs(want).max()) < 1e-13
    # unsorted lines give the same product (the warp segments only merge neighbours)
    perm = rng.permutation(nnz)
    dev2 = H.KRDevice(x[perm], y[perm], z[perm], n)
    v = rng.random(n) + 0.5
    o2 = dev2.vec()
    dev2.spmv(torch.from_numpy(v).to(dev2.device), o2)
    assert np.max(np.abs(o2.cpu().numpy() - A.dot(v)) / np.abs(A.dot(v)).max()) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_bias_matches_reference(lib, name):
    """returnBias on the GPU against the unmodified reference's bias column.  The sums run in another order than scipy's, so
    converging runs agree to ~1e-12; the removed loci and the iteration counts are identical."""
    pytest.importorskip("torch")
    from fithic_b200 import hickry as H
    g = load(name)
    bias = H.returnBias((g["x"], g["y"], g["z"], int(g["n"])), float(g["perc"])).reshape(-1)
    info = H.returnBias.last
    assert np.array_equal(info["removed"], g["removed"])
    ref = g["bias"]
    assert np.array_equal(bias == -1.0, ref == -1.0)
    m = ref > 0
    rel = np.max(np.abs(bias[m] - ref[m]) / ref[m])
    print(name, "outer", info["outer"], int(g["outer"]), "inner", info["inner"], int(g["inner"]), "max rel diff %.2e" % rel)
    if int(g["outer"]) <= 30:  # converged in the reference
        assert (info["outer"], info["inner"]) == (int(g["outer"]), int(g["inner"]))
        assert rel < 1e-9
    else:  # the reference stopped at its iteration cap: an unconverged, ill-conditioned iterate -- same path, looser match
        assert info["outer"] == int(g["outer"])
        assert rel < 1e-4


@pytest.mark.gpu
def test_cli_from_files_to_bias_file(lib, tmp_path, capsys):
    """`HiCKRy.py -i -f -o -x` (HiCKRy.py:264-283) end to end: the bias file a `fithic -t` run would read."""
    pytest.importorskip("torch")
    from fithic_b200 import hickry as H
    g = load("hickry_pfal_10kb")
    ipath, fpath = write_files(g, tmp_path)
    out = str(tmp_path / "bias.gz")
    H.main(["-i", ipath, "-f", fpath, "-o", out, "-x", str(float(g["perc"]))])
    rows = [line.split("\t") for line in gzip.open(out, "rt").read().splitlines()]
    assert [(r[0], int(r[1])) for r in rows] == list(zip(g["chroms"].tolist(), g["mids"].tolist()))
    got, ref = np.array([float(r[2]) for r in rows]), g["bias"]
    assert np.array_equal(got == -1.0, ref == -1.0)
    m = ref > 0
    assert np.max(np.abs(got[m] - ref[m]) / ref[m]) < 1e-9
    text = capsys.readouterr().out
    assert "Creating sparse matrix..." in text and "WARNING" not in text  # mean and median inside (0.5, 2): no warning (:243-250)
