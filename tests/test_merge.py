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
