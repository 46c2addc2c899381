"""Golden fixtures for the KR bias computation from the UNMODIFIED reference fithic/utils/HiCKRy.py on its own bundled
data (every k-th contact line, so that the fixtures stay small).

    python tests/golden/make_golden_hickry.py

Stores the inputs as arrays (locus index pairs in fragment-file order + counts) and the reference's bias column,
including the -1 of the removed loci, plus how many outer / inner iterations its Knight-Ruiz loop took."""
import gzip
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/fithic/tests/data"
sys.path.insert(0, "/root/reference/fithic/utils")

CASES = {  # name: (data set, keep every k-th line, percentOfSparseToRemove)
    "hickry_pfal_10kb": ("Ay_Rings_MboI_Pfal_w10000", 8, 0.05),
    "hickry_hesc_40kb": ("Dixon_hESC_HindIII_hg18_w40000_chr1", 4, 0.05),
    "hickry_pfal_10kb_x10": ("Ay_Rings_MboI_Pfal_w10000", 20, 0.10),
}


def main():
    import HiCKRy as H  # the reference, unmodified
    with tempfile.TemporaryDirectory() as tmp:
        for name, (ds, k, perc) in CASES.items():
            src = os.path.join(DATA, "contactCounts", ds + ".gz")
            frag = os.path.join(DATA, "fragmentLists", ds + ".gz")
            sub = os.path.join(tmp, name + ".gz")
            with gzip.open(src, "rt") as f, gzip.open(sub, "wt", compresslevel=1) as g:
                for i, line in enumerate(f):
                    if i % k == 0:
                        g.write(line)
            matrix, rev = H.loadfastfithicInteractions(sub, frag)
            # the same loop as returnBias, keeping the iteration counts of knightRuizAlg
            mtx, removed = H.removeZeroDiagonalCSR(matrix.copy(), perc)
            res = H.knightRuizAlg(mtx)
            bias = H.addZeroBiases(removed, H.computeBiasVector(res[0]))
            assert np.array_equal(bias, H.returnBias(matrix, perc), equal_nan=True)
            # inputs as arrays
            index = {cm: i for i, cm in enumerate(rev)}
            xs, ys, zs = [], [], []
            with gzip.open(sub, "rt") as f:
                for line in f:
                    w = line.split()
                    xs.append(index[(w[0], int(w[1]))])
                    ys.append(index[(w[2], int(w[3]))])
                    zs.append(float(w[4]))
            np.savez_compressed(os.path.join(HERE, name + ".npz"), x=np.asarray(xs, dtype=np.int32),
                                y=np.asarray(ys, dtype=np.int32), z=np.asarray(zs, dtype=np.float64), n=len(rev), perc=perc,
                                bias=np.asarray(bias, dtype=np.float64).reshape(-1), removed=np.asarray(removed, dtype=np.int64),
                                outer=int(res[1]), inner=int(res[2]),
                                chroms=np.array([c for c, _ in rev]), mids=np.asarray([m for _, m in rev], dtype=np.int64))
            print(name, "loci", len(rev), "lines", len(xs), "removed", len(removed), "outer", res[1], "inner", res[2])


if __name__ == "__main__":
    main()
