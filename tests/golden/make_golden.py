"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (ay-lab/fithic at /root/reference)
in the build container through oracle/ref_harness.py.

    python tests/golden/make_golden.py

Each <case>.npz holds the synthetic inputs (arrays, as written to the gz TSV files the reference read) and, per spline
pass, what the reference computed: N, T, the equal-occupancy bins, x/y, the spline table, full-precision p and q
(captured at myStats.benjamini_hochberg_correction, before the %e formatting of the output file) and the outlier sets.
kat.json holds known answers of scipy.special.bdtrc and of the reference's benjamini_hochberg_correction.
The GPU box has no /root/reference: tests only read these files.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fithic_b200 import synth  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

CASES = {
    # name: (generator kwargs, reference CLI flags after -i/-f/-o/-r)
    "intra_40kb": (dict(n_pairs=30000, res=40000, seed=2001, chroms=["chr21", "chr22"], mean_count=8.0),
                   ["-b", "50"]),
    "intra_bias_LU_p2": (dict(n_pairs=40000, res=40000, seed=2002, chroms=["chr20", "chr21"], mean_count=6.0,
                              with_bias=True), ["-b", "50", "-L", "80000", "-U", "20000000", "-p", "2"]),
    "all_bias": (dict(n_pairs=30000, res=100000, seed=2003, chroms=["chr19", "chr20", "chr21", "chr22"], mean_count=4.0,
                      with_bias=True, inter_fraction=0.4), ["-b", "40", "-x", "All"]),
    "inter_only_bias": (dict(n_pairs=20000, res=100000, seed=2004, chroms=["chr19", "chr20", "chr21", "chr22"],
                             mean_count=3.0, with_bias=True, inter_fraction=1.5), ["-x", "interOnly"]),
    "intra_p3": (dict(n_pairs=30000, res=20000, seed=2005, chroms=["chr22"], mean_count=5.0), ["-b", "100", "-p", "3"]),
}


def run_case(name, gen, flags, tmp):
    contacts, frags, biases, raw = synth.make_intra(**gen)
    res = gen["res"]
    cpath, fpath, bpath = synth.write_inputs(tmp, contacts, frags, res, raw, biases, prefix=name)
    argv = ["-i", cpath, "-f", fpath, "-o", os.path.join(tmp, name + "_out"), "-r", res, "-l", name] + flags
    if bpath:
        argv += ["-t", bpath]
    passes = R.run_reference(argv)
    out = dict(mid1=contacts.mid1, mid2=contacts.mid2, cnt=contacts.cnt, chrs=contacts.chrs,
               chroms=np.array(contacts.chroms), res=res, flags=np.array([str(f) for f in flags]),
               frag_n=frags.n_mappable, frag_maxmid=frags.max_mid, npasses=len(passes))
    if biases is not None:
        out.update(bias_values=biases.values, bias_mids=biases.mids, bias_chr_off=biases.chr_off, bias_raw=raw)
    for i, p in enumerate(passes):
        pre = "p%d_" % (i + 1)
        out[pre + "N"] = p["N"]
        out[pre + "T"] = p["T"]
        out[pre + "scalars"] = np.array([p["observedInterAllCount"], p["observedInterAllSum"], p["observedIntraAllSum"],
                                         p["possibleIntraInRangeCount"]], dtype=np.int64)
        out[pre + "possibleInterAllCount"] = float(p["possibleInterAllCount"])
        out[pre + "interChrProb"] = float(p["interChrProb"])
        md = sorted(p["mainDic"].items())
        out[pre + "dists"] = np.array([d for d, _ in md], dtype=np.int64)
        out[pre + "sums"] = np.array([s for _, s in md], dtype=np.int64)
        b = p["bins"]
        out[pre + "bin_lb"] = np.array([x["lb"] for x in b], dtype=np.int64)
        out[pre + "bin_ub"] = np.array([x["ub"] for x in b], dtype=np.int64)
        out[pre + "bin_pairs"] = np.array([x["pairs"] for x in b], dtype=np.int64)
        out[pre + "bin_sumcc"] = np.array([x["sumcc"] for x in b], dtype=np.int64)
        out[pre + "bin_sumdist"] = np.array([x["sumdist"] for x in b], dtype=np.float64)
        out[pre + "x"] = np.array(p["x"], dtype=np.float64)
        out[pre + "y"] = np.array(p["y"], dtype=np.float64)
        if p["splineX"] is not None:
            out[pre + "splineX"] = np.array(p["splineX"], dtype=np.int64)
            out[pre + "newSplineY"] = np.array(p["newSplineY"], dtype=np.float64)
        out[pre + "p"] = np.array(p["p"], dtype=np.float64)
        out[pre + "q"] = np.array(p["q"], dtype=np.float64)
        out[pre + "outliersline"] = np.array(p["outliersline"], dtype=np.int64)
        out[pre + "outliersdist"] = np.array(p["outliersdist"], dtype=np.int64)
    # first rows of the reference's output file of the last pass (format fixture)
    import gzip
    sig = os.path.join(tmp, name + "_out", "%s.spline_pass%d.res%d.significances.txt.gz" % (name, len(passes), res))
    with gzip.open(sig, "rt") as f:
        lines = f.readlines()
    out["sig_head"] = np.array(lines[:200])
    out["sig_nrows"] = len(lines) - 1
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "passes", len(passes), "lines", len(contacts), "N", [p["N"] for p in passes], "T", passes[0]["T"],
          "outliers", [len(p["outliersline"]) for p in passes])


def make_kat():
    import scipy.special as sc
    F = R.load_reference()
    rng = np.random.default_rng(7)
    bd = []
    for N in (100, 171, 5000, 4219169, 300_000_000, 900_000_000, 2 ** 31 - 1):
        for _ in range(60):
            c = int(min(N, np.floor(np.exp(rng.uniform(0, np.log(min(N, 5000) + 1))))))
            ratio = float(np.exp(rng.uniform(np.log(0.02), np.log(50))))
            prior = min(c * ratio / N, 1.0)
            bd.append((c - 1, N, prior, float(sc.bdtrc(c - 1, N, prior))))
    for k, N, p in ((0, 1000, 1e-3), (0, 1000, 0.5), (-1, 10, 0.3), (10, 10, 0.5), (5, 100, -0.1), (5, 100, 0.0),
                    (5, 100, 1.0), (4, 4219169, 2.4e-06), (49, 4219169, 7.4e-06), (399, 4219169, 7.4e-05),
                    (2, 10 ** 9, 1e-09), (1999, 10 ** 9, 1.5e-06), (0, 5000, 0.0074), (11, 10, 0.5)):
        bd.append((k, N, p, float(sc.bdtrc(k, N, p))))
    bh = []
    nan = float("nan")
    for p, T in (([0.03, 0.4, 0.7, 0.01], 10), ([0.03, 0.4, 0.7, 0.01], 4), ([0.02, 0.02, 1.0, 0.5, 0.02, 1.0, 0.0], 7),
                 ([0.2, nan, 0.01, 1.0, nan, 0.9], 6), ([0.01, 0.011, 0.012, 0.5], 4), ([1e-9, 1e-3, 0.5], 1000)):
        bh.append((p, T, [float(v) for v in F.myStats.benjamini_hochberg_correction(list(p), T)]))
    for n in (50, 1000):
        p = (rng.random(n) ** 3).tolist()
        for i in rng.integers(0, n, n // 5):
            p[int(i)] = 1.0
        for i in rng.integers(0, n, n // 10):
            p[int(i)] = p[int(rng.integers(0, n))]
        T = int(n * 3.7)
        bh.append((p, T, [float(v) for v in F.myStats.benjamini_hochberg_correction(list(p), T)]))
    import scipy
    import sklearn
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump({"generated_with": {"scipy": scipy.__version__, "sklearn": sklearn.__version__,
                                      "numpy": np.__version__, "reference": "ay-lab/fithic 2.0.7 (/root/reference)"},
                   "bdtrc": bd, "bh": bh}, f)
    print("kat.json: %d bdtrc, %d bh vectors" % (len(bd), len(bh)))


if __name__ == "__main__":
    if not R.reference_available():
        raise SystemExit("needs /root/reference (build container)")
    with tempfile.TemporaryDirectory() as tmp:
        for name, (gen, flags) in CASES.items():
            run_case(name, gen, flags, tmp)
    make_kat()
