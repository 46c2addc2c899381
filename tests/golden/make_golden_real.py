"""Golden fixtures from the reference's OWN bundled data sets (subsampled so that the fixtures stay small).

    python tests/golden/make_golden_real.py

Takes every k-th line of a contact file of /root/reference/fithic/tests/data (plus its full fragment / bias files), runs
the UNMODIFIED reference on it through oracle/ref_harness.py and stores inputs + per-pass results like make_golden.py.
Cases follow fithic/tests/run_tests-git.sh:34-54 (flags in CASES below)."""
import gzip
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fithic_b200 import io as fio  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

DATA = "/root/reference/fithic/tests/data"
CASES = {
    # name: (data set, resolution, keep every k-th line, flags)  -- run_tests-git.sh:52-54, :40-42
    "real_pfal_10kb": ("Ay_Rings_MboI_Pfal_w10000", 10000, 25, ["-b", "200", "-x", "All"]),
    "real_hesc_40kb_bias": ("Dixon_hESC_HindIII_hg18_w40000_chr1", 40000, 12,
                            ["-L", "50000", "-U", "5000000", "-b", "50", "-x", "intraOnly", "-p", "2", "-t", "BIAS"]),
    # restriction-fragment mode (-r 0), run_tests-git.sh:28-30, two passes instead of one to cover the outlier bookkeeping
    "real_hesc_refrags_r0": ("Dixon_hESC_HindIII_hg18_combineFrags10_chr1", 0, 12,
                             ["-L", "50000", "-U", "5000000", "-b", "200", "-x", "intraOnly", "-p", "2"]),
    # the same with a bias file: there is none for the restriction fragments among the bundled data, so the reference's own
    # HiCKRy (unmodified) computes one from the sub-sampled contacts first
    "real_mesc_refrags_r0_bias": ("Dixon_mESC_HindIII_mm9_combineFrags10_chr1", 0, 16,
                                  ["-L", "50000", "-U", "5000000", "-b", "100", "-x", "intraOnly", "-p", "1", "-t", "KRBIAS"]),
}


def main():
    if not R.reference_available():
        raise SystemExit("needs /root/reference")
    with tempfile.TemporaryDirectory() as tmp:
        for name, (ds, res, k, flags) in CASES.items():
            if len(sys.argv) > 1 and name not in sys.argv[1:]:
                continue
            src = os.path.join(DATA, "contactCounts", ds + ".gz")
            sub = os.path.join(tmp, name + ".contacts.gz")
            with gzip.open(src, "rt") as f, gzip.open(sub, "wt", compresslevel=1) as g:
                for i, line in enumerate(f):
                    if i % k == 0:
                        g.write(line)
            frag = os.path.join(DATA, "fragmentLists", ds + ".gz")
            bias = os.path.join(DATA, "biasPerLocus", ds + ".gz")
            if "KRBIAS" in flags:
                sys.path.insert(0, "/root/reference/fithic/utils")
                import HiCKRy as H  # the reference's bias generator, unmodified
                bias = os.path.join(tmp, name + ".bias.gz")
                matrix, rev = H.loadfastfithicInteractions(sub, frag)
                H.outputBias(H.returnBias(matrix, 0.05), rev, bias)
            fl = [bias if x in ("BIAS", "KRBIAS") else x for x in flags]
            argv = ["-i", sub, "-f", frag, "-o", os.path.join(tmp, name + "_out"), "-r", res, "-l", name] + fl
            passes = R.run_reference(argv)
            contacts = fio.read_contacts(sub)
            chroms = list(contacts.chroms)
            frags = fio.read_fragments(frag, chroms, 1, keep_mids=(res == 0))
            contacts.chroms = chroms
            out = dict(mid1=contacts.mid1, mid2=contacts.mid2, cnt=contacts.cnt, chrs=contacts.chrs,
                       chroms=np.array(chroms), res=res,
                       flags=np.array([str(x) for x in flags if x not in ("-t", "BIAS", "KRBIAS")]),
                       frag_n=frags.n_mappable, frag_maxmid=frags.max_mid, npasses=len(passes))
            if res == 0:  # the fragment mid points themselves, concatenated chromosome by chromosome (frag_n gives the split)
                out["frag_mids"] = np.concatenate([np.asarray(m, dtype=np.int64) for m in frags.mids])
            if "BIAS" in flags or "KRBIAS" in flags:
                b, _ = fio.read_biases(bias, chroms, res, 0.5, 2.0)
                out.update(bias_values=b.values, bias_mids=b.mids, bias_chr_off=b.chr_off)
                out["frag_n"], out["frag_maxmid"] = frags.n_mappable, frags.max_mid
                out["chroms"] = np.array(chroms)
            for i, p in enumerate(passes):
                pre = "p%d_" % (i + 1)
                out[pre + "N"], out[pre + "T"] = p["N"], p["T"]
                out[pre + "scalars"] = np.array([p["observedInterAllCount"], p["observedInterAllSum"],
                                                 p["observedIntraAllSum"], p["possibleIntraInRangeCount"]], dtype=np.int64)
                out[pre + "possibleInterAllCount"] = float(p["possibleInterAllCount"])
                out[pre + "interChrProb"] = float(p["interChrProb"])
                md = sorted(p["mainDic"].items())
                out[pre + "dists"] = np.array([d for d, _ in md], dtype=np.int64)
                out[pre + "sums"] = np.array([s for _, s in md], dtype=np.int64)
                bn = p["bins"]
                for key, dt in (("lb", np.int64), ("ub", np.int64), ("pairs", np.int64), ("pairs7", np.int64),
                                ("sumcc", np.int64), ("sumdist", np.float64)):
                    out[pre + "bin_" + key] = np.array([x[key] for x in bn], dtype=dt)
                out[pre + "x"] = np.array(p["x"], dtype=np.float64)
                out[pre + "y"] = np.array(p["y"], dtype=np.float64)
                if p["splineX"] is not None:
                    out[pre + "splineX"] = np.array(p["splineX"], dtype=np.int64)
                    out[pre + "newSplineY"] = np.array(p["newSplineY"], dtype=np.float64)
                out[pre + "p"] = np.array(p["p"], dtype=np.float64)
                out[pre + "q"] = np.array(p["q"], dtype=np.float64)
                out[pre + "outliersline"] = np.array(p["outliersline"], dtype=np.int64)
                out[pre + "outliersdist"] = np.array(p["outliersdist"], dtype=np.int64)
            sig = os.path.join(tmp, name + "_out", "%s.spline_pass%d%s.significances.txt.gz" %
                               (name, len(passes), ".res%d" % res if res else ""))
            with gzip.open(sig, "rt") as f:
                lines = f.readlines()
            out["sig_head"] = np.array(lines[:200])
            out["sig_nrows"] = len(lines) - 1
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
            print(name, "lines", len(contacts), "passes", len(passes), "N", [p["N"] for p in passes], "T", passes[0]["T"],
                  "outliers", [len(p["outliersline"]) for p in passes])


if __name__ == "__main__":
    main()
