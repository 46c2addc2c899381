"""Golden fixtures for the merge-filter step from the UNMODIFIED reference: fithic/fithic.py produces the significances file
of a bundled data set, the awk line of fithic/utils/merge-filter.sh keeps the rows with q <= fdr, and
fithic/utils/CombineNearbyInteraction.py (networkx as installed here) merges them under several option sets.

    python tests/golden/make_golden_merge.py

Stores the filtered rows as arrays (the p and q columns as the doubles `float(text)` gives) and the complete text of every
output file."""
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
DATA = "/root/reference/fithic/tests/data"
SCRIPT = "/root/reference/fithic/utils/CombineNearbyInteraction.py"

CASES = {  # name: (data set, resolution, extra fithic flags, fdr, {variant: extra merge flags})
    "merge_pfal_10kb": ("Ay_Rings_MboI_Pfal_w10000", 10000, [], 0.01, {
        "default": [], "p50": ["-p", "50"], "p0": ["-p", "0"], "c4": ["-c", "4"], "n1": ["-n", "1"], "n0": ["-n", "0"],
        "s1": ["-s", "1"], "s1_p50": ["-s", "1", "-p", "50"], "s1_p0": ["-s", "1", "-p", "0"], "p7_c4": ["-p", "7", "-c", "4"]}),
    "merge_hesc_40kb": ("Dixon_hESC_HindIII_hg18_w40000_chr1", 40000,
                        ["-t", DATA + "/biasPerLocus/Dixon_hESC_HindIII_hg18_w40000_chr1.gz"], 1e-20, {
        "default": [], "p0": ["-p", "0"], "p30": ["-p", "30"]}),
}


def synthetic_rows(seed=7, n=1800, res=5000):
    """Rows no bundled data set has: repeated and mirrored bin pairs, inter-chromosomal lines, a chromosome that occurs in
    column 1 only on inter lines, many ties in q and in the count."""
    rng = np.random.default_rng(seed)
    names = ["chrB", "chrA", "chr10", "chrZ"]
    rows = []
    for _ in range(n):
        c1 = names[int(rng.integers(0, 3))]
        c2 = c1 if rng.random() < 0.9 else names[int(rng.integers(0, 3))]
        if rng.random() < 0.02:
            c1, c2 = "chrZ", "chrA"
        a = int(rng.integers(0, 45))
        b = min(44, a + int(rng.integers(0, 6))) if rng.random() < 0.8 else int(rng.integers(0, 45))
        if rng.random() < 0.2:
            a, b = b, a
        q = float(rng.choice([1e-30, 2.5e-12, 1e-7, 3e-4, 0.004]))
        pv = q * float(rng.choice([0.01, 0.1, 0.1, 0.5]))
        rows.append([c1, str(a * res + res // 2), c2, str(b * res + res // 2), str(int(rng.integers(1, 6))), "%e" % pv,
                     "%e" % q, "1.000000e+00", "1.000000e+00", "1.000000"])
    return rows, res


def store_case(tmp, name, sub, res, fdr, variants):
    rows = [line.split() for line in gzip.open(sub, "rt")]
    store = dict(res=res, fdr=fdr, chr1=np.array([r[0] for r in rows]), chr2=np.array([r[2] for r in rows]),
                 mid1=np.array([int(r[1]) for r in rows], dtype=np.int64),
                 mid2=np.array([int(r[3]) for r in rows], dtype=np.int64),
                 cc=np.array([int(r[4]) for r in rows], dtype=np.int64),
                 p=np.array([float(r[5]) for r in rows], dtype=np.float64),
                 q=np.array([float(r[6]) for r in rows], dtype=np.float64),
                 variants=np.array(sorted(variants)))
    for v, flags in variants.items():
        merged = os.path.join(tmp, name + "_" + v, "merged.gz")
        src = sub
        if "-H" in flags:  # a file with a header line, as fithic writes it (the default of the script, :86)
            src = os.path.join(tmp, name + "_with_header.gz")
            with gzip.open(src, "wt") as f:
                f.write("chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\tbias1\tbias2\tExpCC\n")
                f.write(gzip.open(sub, "rt").read())
        head = [] if "-H" in flags else ["-H", "0"]  # merge-filter.sh:23
        subprocess.run([sys.executable, SCRIPT, "-i", src] + head + ["-r", str(res), "-o", merged] + flags,
                       check=True, stdout=subprocess.DEVNULL)
        text = gzip.open(merged, "rt").read()
        store["out_" + v] = np.frombuffer(text.encode(), dtype=np.uint8)
        store["flags_" + v] = np.array(flags if flags else [""])
        print(name, v, "rows in", len(rows), "rows out", text.count("\n"))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


def main():
    with tempfile.TemporaryDirectory() as tmp:
        rows, res = synthetic_rows()
        sub = os.path.join(tmp, "synth_subset.gz")
        with gzip.open(sub, "wt") as f:
            f.write("".join("\t".join(r) + "\n" for r in rows))
        store_case(tmp, "merge_synth", sub, res, 1.0, {
            "default": [], "p0": ["-p", "0"], "p40": ["-p", "40"], "c4_n1": ["-c", "4", "-n", "1"], "s1": ["-s", "1"],
            "s1_p0": ["-s", "1", "-p", "0"], "H1": ["-H", "1"], "c6": ["-c", "6"], "p150": ["-p", "150"]})
    if "--synthetic-only" in sys.argv:
        return
    from oracle import ref_harness as R
    F = R.load_reference()
    with tempfile.TemporaryDirectory() as tmp:
        for name, (ds, res, extra, fdr, variants) in CASES.items():
            out = os.path.join(tmp, name)
            argv = sys.argv
            sys.argv = ["fithic", "-i", DATA + "/contactCounts/" + ds + ".gz", "-f", DATA + "/fragmentLists/" + ds + ".gz", "-o",
                        out, "-r", str(res), "-l", "lib", "-p", "1", "-x", "intraOnly"] + extra
            with open(os.devnull, "w") as null:
                stdout, sys.stdout = sys.stdout, null
                try:
                    F.main()
                finally:
                    sys.stdout, sys.argv = stdout, argv
            sig = os.path.join(out, "lib.spline_pass1.res%d.significances.txt.gz" % res)
            sub = os.path.join(tmp, name + "_subset.gz")
            # merge-filter.sh:22
            subprocess.run("zcat %s | awk '{if(NR!=1){print $0}}' | awk -v q=\"%s\" '{if($7<=q){print $0}}' | gzip > %s"
                           % (sig, fdr, sub), shell=True, check=True)
            store_case(tmp, name, sub, res, fdr, variants)


if __name__ == "__main__":
    main()
