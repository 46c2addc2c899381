"""Merge-filter step: oracle/merge_oracle.py against output files captured from the unmodified reference
fithic/utils/CombineNearbyInteraction.py (CPU), and fithic_b200/merge.py on the GPU against the same files, byte for byte."""
import gzip
import os

import numpy as np
import pytest

from oracle import merge_oracle as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["merge_synth", "merge_pfal_10kb", "merge_hesc_40kb"]
FLAG_NAMES = {"-p": "top_pct", "-c": "conn", "-n": "neigh", "-s": "sort_order", "-H": "header"}


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def variants(name):
    z = load(name)
    return [(name, str(v)) for v in z["variants"]]


ALL = [nv for c in CASES for nv in variants(c)]


def options(z, variant):
    flags = [f for f in z["flags_" + variant].tolist() if f]
    return {FLAG_NAMES[flags[i]]: int(flags[i + 1]) for i in range(0, len(flags), 2)}


def write_rows(z, path, header):
    """The fixture's rows as a significances file (p and q as the shortest text that reads back to the same double)."""
    with gzip.open(path, "wt", compresslevel=1) as f:
        if header:
            f.write("chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\tbias1\tbias2\tExpCC\n")
        f.write("".join("%s\t%d\t%s\t%d\t%d\t%r\t%r\t1.0\t1.0\t1.0\n" % row for row in
                        zip(z["chr1"].tolist(), z["mid1"].tolist(), z["chr2"].tolist(), z["mid2"].tolist(), z["cc"].tolist(),
                            z["p"].tolist(), z["q"].tolist())))


@pytest.mark.parametrize("name,variant", ALL, ids=["%s-%s" % nv for nv in ALL])
def test_oracle_reproduces_reference_output(name, variant):
    """Every byte of the reference's output file: chromosome order, component order, the rows kept by the greedy
    neighbourhood filter, the set-order dependent choice of -p 0, the empty result of -s 1 with 0 < -p < 100."""
    z = load(name)
    kw = options(z, variant)
    kw.pop("header", None)
    text = M.merge_rows(z["chr1"].tolist(), z["mid1"].tolist(), z["chr2"].tolist(), z["mid2"].tolist(), z["cc"].tolist(),
                        z["p"].tolist(), z["q"].tolist(), int(z["res"]), **kw)
    assert text == z["out_" + variant].tobytes().decode()


def test_custom_percent_and_bins():
    assert M.custom_percent([5, 1, 3], 50, 1) == 5 and M.custom_percent([5, 1, 3], 50, 2) == 1  # index <= 1: max / min
    assert M.custom_percent(list(range(10)), 50, 1) == 5 and M.custom_percent(list(range(10)), 50, 2) == 4
    assert M.bin_of(20000, 40000) == 1.0 and M.bin_of("60000.0", 40000) == 2.0
    assert M.chromosome_order(["chr2", "chr10", "chr1", "chr2"]) == ["chr1", "chr10", "chr2"]


def rows_of(z):
    return dict(chr1=z["chr1"].astype(object), chr2=z["chr2"].astype(object), mid1=z["mid1"].astype(np.float64),
                mid2=z["mid2"].astype(np.float64), cc=z["cc"].astype(np.int64), p=z["p"], q=z["q"])


@pytest.mark.parametrize("name,variant", ALL, ids=["%s-%s" % nv for nv in ALL])
def test_library_code_on_host_reproduces_reference_output(lib, name, variant):
    """The per-entry functions of csrc/merge.cu (keys, neighbour search, union-find, box rows, heap order, neighbourhood
    rule) driven serially by fhc_host_merge_*, under the host assembly of fithic_b200/merge.py: the reference's file."""
    from fithic_b200 import merge as G
    from tests.util import merge_components_host
    z = load(name)
    kw = options(z, variant)
    kw.pop("header", None)
    text = G.merge_rows(rows_of(z), int(z["res"]), components=merge_components_host, **kw)
    assert text == z["out_" + variant].tobytes().decode()


def test_reader_and_refusals(tmp_path):
    from fithic_b200 import merge as G
    z = load("merge_synth")
    path = str(tmp_path / "sig.gz")
    write_rows(z, path, header=True)
    rows = G.read_rows(path, header=1)
    for k, v in rows_of(z).items():
        assert np.array_equal(rows[k], v), k
    with pytest.raises(ValueError):  # a header line read as data: float("fragmentMid1") fails in the reference as well
        G.read_rows(path, header=0)
    sub = G.read_rows(path, header=1, fdr=1e-7)
    assert len(sub["q"]) == int((z["q"] <= 1e-7).sum()) and (sub["q"] <= 1e-7).all()
    with pytest.raises(ValueError):
        G.bins_of(np.array([20001.0]), 5000)  # off the grid: fractional bin number in the reference
    assert G.bins_of(np.array([2500.0, 7500.0]), 5000).tolist() == [1, 2]
    empty = str(tmp_path / "empty.gz")
    with gzip.open(empty, "wt"):
        pass
    assert G.merge_rows(G.read_rows(empty, header=0), 5000, components=lambda *a: None) == G.HEADER
    o = G.parse_args(["-i", "a", "-o", "b", "-r", "5000"])
    assert (o.headerInp, o.connectivity_rule, o.TopPctElem, o.NeighborHoodBin, o.SortOrder) == (1, 8, 100, 2, 0)
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):  # no CPU fallback on the product path
            G.merge_rows(rows, 5000)


def test_merge_filter_command_line_with_fdr(lib, tmp_path, monkeypatch):
    """merge-filter.sh's arguments: header dropped, rows with q <= fdr kept (awk's numeric compare), default options."""
    from fithic_b200 import merge as G
    from tests.util import merge_components_host
    monkeypatch.setattr(G, "components_device", merge_components_host)  # no GPU in the CPU suite
    z = load("merge_synth")
    full, out = str(tmp_path / "sig.gz"), str(tmp_path / "deep" / "dir" / "merged.gz")
    write_rows(z, full, header=True)
    G.main([full, str(int(z["res"])), out, "1e-7"])
    keep = z["q"] <= 1e-7
    want = M.merge_rows(z["chr1"][keep].tolist(), z["mid1"][keep].tolist(), z["chr2"][keep].tolist(), z["mid2"][keep].tolist(),
                        z["cc"][keep].tolist(), z["p"][keep].tolist(), z["q"][keep].tolist(), int(z["res"]))
    assert gzip.open(out, "rt").read() == want and want.count("\n") > 10
    with pytest.raises(SystemExit):
        G.main([full, "5000"])


def test_union_find_on_long_chains_and_blocks(lib):
    """Shapes the data sets do not have: one 20000-node diagonal chain (a deep union-find tree without path halving), a
    filled block, isolated nodes, a repeated pair; components, boxes and the box census against a scipy labelling."""
    import scipy.ndimage as ndi
    from tests.util import merge_components_host
    rng = np.random.default_rng(3)
    b1 = list(range(1, 20001)) + [5 + i for i in range(30) for _ in range(30)] + [100, 100, 300]
    b2 = list(range(2, 20002)) + [25000 + j for _ in range(30) for j in range(30)] + [9000, 9000, 9500]
    perm = rng.permutation(len(b1))
    b1, b2 = np.asarray(b1)[perm], np.asarray(b2)[perm]
    cc = rng.integers(1, 9, len(b1))
    g = merge_components_host(np.zeros(len(b1), np.int32), b1, b2, cc, rng.random(len(b1)), 8, 100, 2, 0)
    roots = np.nonzero(g["label"] == np.arange(len(b1)))[0]
    assert sorted(g["size"][roots].tolist()) == [1, 1, 900, 20000]
    assert (g["label"] == -1).sum() == 1
    big = roots[np.argmax(g["size"][roots])]
    # the box of the chain also holds the two isolated nodes (other components count, :389-394); the block lies outside
    assert g["box"][big].tolist() == [1, 20000, 2, 20001] and g["have"][big] == 20000 + 2
    block = roots[g["size"][roots] == 900][0]
    assert g["box"][block].tolist() == [5, 34, 25000, 25029] and g["have"][block] == 900
    # dense random field against scipy's 8-connected labelling
    field = rng.random((60, 60)) < 0.35
    field = np.triu(field)
    ys, xs = np.nonzero(field)
    g = merge_components_host(np.zeros(len(ys), np.int32), ys + 1, xs + 1, np.ones(len(ys), np.int64), rng.random(len(ys)), 8,
                              100, 2, 0)
    lab, ncomp = ndi.label(field, structure=np.ones((3, 3)))
    assert (g["label"] == np.arange(len(ys))).sum() == ncomp
    by_line = np.empty(len(ys), np.int64)
    by_line[g["order"]] = g["label"]  # label of the entry each input line became
    assert len(set(zip(lab[ys, xs].tolist(), by_line.tolist()))) == ncomp  # the two labellings are the same partition


def random_case(seed, nmax):
    """Random rows on a small grid (many ties, repeated and mirrored pairs, inter lines, q = 0) and a random option set,
    legal or not (-c 5 adds no edges, -p 101 prints nothing, -n -1 keeps everything)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, nmax))
    res = int(rng.choice([2, 1000, 5000, 40000]))
    nb = int(rng.integers(3, 60))
    names = ["chr" + str(i) for i in rng.permutation(12)[: int(rng.integers(1, 5))]]
    c1 = rng.choice(names, n)
    c2 = np.where(rng.random(n) < 0.9, c1, rng.choice(names, n))
    a = rng.integers(0, nb, n)
    b = np.clip(a + rng.integers(-3, 8, n), 0, nb - 1)
    q = rng.choice([0.0, 1e-30, 2.5e-12, 1e-7, 3e-4, 0.004, 0.5], n)
    rows = dict(chr1=c1.astype(object), chr2=c2.astype(object), mid1=(a * res + res // 2).astype(np.float64),
                mid2=(b * res + res // 2).astype(np.float64), cc=rng.integers(0, 5, n).astype(np.int64),
                p=q * rng.choice([0.01, 0.1, 1.0], n), q=q)
    kw = dict(conn=int(rng.choice([8, 8, 4, 5])), top_pct=int(rng.choice([100, 100, 0, 1, 25, 50, 99, 101, -3])),
              neigh=int(rng.choice([2, 0, 1, 3, -1])), sort_order=int(rng.choice([0, 0, 1])))
    return rows, res, kw


def oracle_text(rows, res, kw):
    return M.merge_rows(rows["chr1"].tolist(), rows["mid1"].tolist(), rows["chr2"].tolist(), rows["mid2"].tolist(),
                        rows["cc"].tolist(), rows["p"].tolist(), rows["q"].tolist(), res, **kw)


def test_library_code_on_host_against_oracle_on_random_rows(lib):
    from fithic_b200 import merge as G
    from tests.util import merge_components_host
    for seed in range(60):
        rows, res, kw = random_case(seed, 1200)
        assert G.merge_rows(rows, res, components=merge_components_host, **kw) == oracle_text(rows, res, kw), (seed, kw)


@pytest.mark.skipif(not os.path.isfile("/root/reference/fithic/utils/CombineNearbyInteraction.py"),
                    reason="needs the reference checkout (build container only)")
def test_oracle_against_the_reference_script_on_random_rows(tmp_path):
    """Container only: the unmodified script run on random rows and option sets (120 further seeds were run once by hand)."""
    import subprocess
    import sys
    for seed in range(1000, 1006):
        rows, res, kw = random_case(seed, 300)
        src = str(tmp_path / ("in%d.gz" % seed))
        with gzip.open(src, "wt") as f:
            f.write("".join("%s\t%d\t%s\t%d\t%d\t%r\t%r\t1\t1\t1\n" % r for r in
                            zip(rows["chr1"].tolist(), rows["mid1"].astype(np.int64).tolist(), rows["chr2"].tolist(),
                                rows["mid2"].astype(np.int64).tolist(), rows["cc"].tolist(), rows["p"].tolist(),
                                rows["q"].tolist())))
        out = str(tmp_path / ("o%d" % seed) / "m.gz")
        subprocess.run([sys.executable, "/root/reference/fithic/utils/CombineNearbyInteraction.py", "-i", src, "-H", "0", "-r",
                        str(res), "-o", out, "-c", str(kw["conn"]), "-p", str(kw["top_pct"]), "-n", str(kw["neigh"]), "-s",
                        str(kw["sort_order"])], check=True, stdout=subprocess.DEVNULL)
        assert gzip.open(out, "rt").read() == oracle_text(rows, res, kw), (seed, kw)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,variant", ALL, ids=["%s-%s" % nv for nv in ALL])
def test_gpu_reproduces_reference_output(lib, name, variant):
    """fithic_b200.merge on the GPU: every byte of the unmodified reference's output file."""
    from fithic_b200 import merge as G
    z = load(name)
    kw = options(z, variant)
    kw.pop("header", None)
    assert G.merge_rows(rows_of(z), int(z["res"]), **kw) == z["out_" + variant].tobytes().decode()


@pytest.mark.gpu
def test_gpu_matches_host_drivers_on_a_large_random_field(lib):
    """200k lines on 3 chromosomes (dense enough for components of thousands of nodes): the kernels' arrays equal the serial
    run of the same code, whatever order the atomics came in."""
    from fithic_b200 import merge as G
    from tests.util import merge_components_host
    rng = np.random.default_rng(11)
    n = 200_000
    chr_rank = rng.integers(0, 3, n).astype(np.int32)
    a = rng.integers(1, 700, n)
    b = np.minimum(a + rng.integers(0, 40, n), 740)
    cc = rng.integers(1, 50, n)
    q = rng.choice([1e-30, 1e-12, 3e-9, 1e-5, 0.01], n) * rng.choice([1.0, 0.5], n)
    for conn, top, neigh, so in [(8, 100, 2, 0), (4, 37, 1, 0), (8, 100, 3, 1)]:
        h = merge_components_host(chr_rank, a, b, cc, q, conn, top, neigh, so)
        g = G.components_device(chr_rank, a, b, cc, q, conn, top, neigh, so)
        for k in h:
            assert np.array_equal(np.asarray(g[k]), np.asarray(h[k])), (k, conn, top)


@pytest.mark.gpu
def test_gpu_cli_and_merge_filter_on_files(lib, tmp_path):
    """`CombineNearbyInteraction.py -i -H 0 -r -o` and merge-filter.sh on files."""
    from fithic_b200 import merge as G
    z = load("merge_synth")
    sub, full = str(tmp_path / "subset.gz"), str(tmp_path / "sig.gz")
    write_rows(z, sub, header=False)
    write_rows(z, full, header=True)
    out = str(tmp_path / "o" / "merged.gz")
    G.main(["-i", sub, "-H", "0", "-r", str(int(z["res"])), "-o", out])
    assert gzip.open(out, "rt").read() == z["out_default"].tobytes().decode()
    G.main(["-i", full, "-r", str(int(z["res"])), "-o", out, "-c", "4", "-n", "1"])
    assert gzip.open(out, "rt").read() == z["out_c4_n1"].tobytes().decode()
    G.merge_filter(full, int(z["res"]), out, 1.0)
    assert gzip.open(out, "rt").read() == z["out_default"].tobytes().decode()
    os.remove(out)
    G.main([full, str(int(z["res"])), out, "1.0", "unused/utility/folder/"])  # merge-filter.sh's positional arguments
    assert gzip.open(out, "rt").read() == z["out_default"].tobytes().decode()
