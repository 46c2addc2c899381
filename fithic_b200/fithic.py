"""`fithic` command line on the B200 path -- same flags, file names and output format as the reference CLI
(fithic/fithic.py:43-124 parse_args, :129-379 main).

    fithic -i CONTACTS.gz -f FRAGS.gz -o OUTDIR -r RES [-t BIAS.gz] [-p N] [-b N] [-m N] [-l LIB] [-U bp] [-L bp]
           [-x intraOnly|interOnly|All] [-tL f] [-tU f] [-V]

`-r 0` (restriction-fragment mode) is supported; `-v` (plots) is outside the accelerated path and ignored with a message;
everything numeric is computed on the GPU (no CPU fallback).

Several GPUs of one box:  torchrun --nproc-per-node N -m fithic_b200 <the same flags>
Every rank scores a contiguous slice of the contact file (cut at chromosome boundaries where one is near: "contacts shard
by chromosome"), the distance histogram and the BH ranks are global (fithic_b200/parallel.py), every rank writes the gzip
members of its own rows and rank 0 joins them in file order -- the output file has the same rows as a single-GPU run.
"""
import argparse
import os
import sys
import time

import numpy as np

from . import __version__
from . import io as fio
from .engine import Engine, Settings


def write_log(logfile, r, st, frags, bias_log, table_path):
    """${lib}.fithic.log with the reference's sections and wording (re-opened 'w' in every pass like the reference,
    fithic/fithic.py:444; sections :464-468, :550-552, :563-569, :648-650 / :737-739, :782-791, :844-846, :915-917,
    :926-928, :1228-1231).  The bias lines (:812-815, :834) sit behind the fragment section in the first pass; a later pass
    re-opens the log and does not read the biases again, so its log has none -- as in the reference."""
    res, L, U = st.resolution, st.L, st.U
    dists = r.get("dists")
    with open(logfile, "w") as log:
        log.write("\n\nInteractions file read successfully\n")
        log.write("------------------------------------------------------------------------------------\n")
        log.write("Observed, Intra-chr in range: pairs= %d\t totalCount= %d\n" % (r["observedIntraInRangeLines"], r["N"]))
        log.write("Observed, Intra-chr all: pairs= %d\t totalCount= %d\n" % (r["observedIntraAllLines"], r["observedIntraAllSum"]))
        log.write("Observed, Inter-chr all: pairs= %d\t totalCount= %d\n" % (r["observedInterAllCount"], r["observedInterAllSum"]))
        if dists is not None and len(dists):
            log.write("Range of observed genomic distances [%s %s]\n" % (int(dists[0]), int(dists[-1])))
        else:
            log.write("Range of observed genomic distances [%s %s]\n" % (float("inf"), 0))
        log.write("\n")
        log.write("Making equal occupancy bins\n")
        log.write("------------------------------------------------------------------------------------\n")
        log.write("Observed intra-chr read counts in range\t%r\nDesired number of contacts per bin\t%r,\nNumber of bins\t%r\n"
                  % (r["N"], r["N"] / st.noOfBins, st.noOfBins))
        log.write("Equal occupancy bins generated\n")
        log.write("\n")
        log.write(("Looping through" if res else "Enumerating") + " all possible fragment pairs in-range\n")
        log.write("------------------------------------------------------------------------------------\n")
        order = [i for i in sorted(range(len(frags.chroms)), key=lambda i: frags.chroms[i]) if frags.n_mappable[i] > 0]
        noOfFrags = int(sum(int(frags.n_mappable[i]) for i in order))
        has_bins = r["bins"]["n"] > 0
        min_possible = float("inf")
        for i in order:
            n = int(frags.n_mappable[i])
            if res:  # :613-643: npairs = n - k per distance step in range, counted twice when bins exist
                stop = int(float(frags.max_mid[i]) - res / 2.0 + 1.0)
                nsteps = (stop + res - 1) // res if stop > 0 else 0
                k0 = 0 if L <= 0 else (L + res - 1) // res
                k1 = nsteps - 1 if U < 0 else min(nsteps - 1, U // res)
                per_chr = 0
                if k1 >= k0:
                    cnt = k1 - k0 + 1
                    per_chr = (n * cnt - (k0 + k1) * cnt // 2) * (2 if has_bins else 1)
                    min_possible = min(min_possible, k0 * res)
            else:
                from .engine import varsize_pairs_in_range
                f = np.sort(np.asarray(frags.mids[i], dtype=np.int64))
                per_chr = varsize_pairs_in_range(f, np.array([0, len(f)], dtype=np.int64), L, U)
                if per_chr:  # the closest pair in range (:713)
                    idx = np.arange(len(f), dtype=np.int64)
                    lo = np.maximum(np.searchsorted(f, f + max(L, 0), side="left"), idx + 1)
                    ok = lo < len(f)
                    d = f[lo[ok]] - f[ok]
                    d = d[d <= U] if U >= 0 else d
                    if len(d):
                        min_possible = min(min_possible, int(d.min()))
            log.write("Chromosome %r,\t%d mappable fragments, \t%d possible intra-chr fragment pairs in range,\t%d possible "
                      "inter-chr fragment pairs\n" % (frags.chroms[i], n, per_chr, (noOfFrags - n) * n))
        log.write("Number of all fragments= %s\n" % noOfFrags)
        log.write("Possible, Intra-chr in range: pairs= %s \n" % r["possibleIntraInRangeCount"])
        log.write("Possible, Intra-chr all: pairs= %s \n" % r["possibleIntraAllCount"])
        log.write("Possible, Inter-chr all: pairs= %s \n" % r["possibleInterAllCount"])
        log.write("Desired genomic distance range   [%d %s] \n" % (L if L > 0 else 0, U if U >= 0 else float("inf")))
        if res:  # :604: the largest maxFrag = max(mid) - res / 2 over the chromosomes
            max_possible = max([float(frags.max_mid[i]) - res / 2.0 for i in order] or [0])
        else:
            max_possible = r.get("maxPossibleGenomicDist", 0)
        log.write("Range of possible genomic distances  [%s  %d] \n"
                  % ("%d" % min_possible if min_possible != float("inf") else "inf", max_possible))
        pia = r["possibleIntraAllCount"]
        log.write("Baseline intrachromosomal probability is %s \n" % (1.0 / pia if pia > 0 else 0))
        log.write("Interchromosomal probability is %s \n" % (r["interChrProb"] if r["interChrProb"] else 0))
        for line in bias_log:
            log.write(line + "\n")
        if bias_log:
            log.write("\n")
        log.write("\nCalculating probability means and standard deviations of contact counts\n")
        log.write("------------------------------------------------------------------------------------\n")
        log.write("Means and error written to %s\n" % table_path)
        log.write("\n")
        log.write("\nFitting a univariate spline to the probability means\n")
        log.write("------------------------------------------------------------------------------------\n")
        log.write("Spline successfully fit\n")
        log.write("\n")
        log.write("\n")


def parse_args(args):
    parser = argparse.ArgumentParser(description="Check the help flag")
    parser.add_argument("-i", "--interactions", dest="intersfile", required=True,
                        help="REQUIRED: interactions between fragment pairs are read from INTERSFILE")
    parser.add_argument("-f", "--fragments", dest="fragsfile", required=True,
                        help="REQUIRED: midpoints (or start indices) of the fragments are read from FRAGSFILE")
    parser.add_argument("-o", "--outdir", dest="outdir", required=True,
                        help="REQUIRED: where the output files will be written")
    parser.add_argument("-r", "--resolution", dest="resolution", type=int, required=True,
                        help="REQUIRED: resolution of the fixed-size dataset; 0 = restriction-fragment (non fixed size) data")
    parser.add_argument("-t", "--biases", dest="biasfile", required=False,
                        help="RECOMMENDED: biases calculated by ICE or KR norm for each locus are read from BIASFILE")
    parser.add_argument("-p", "--passes", dest="noOfPasses", type=int, required=False,
                        help="OPTIONAL: number of spline passes to run. Default is 1")
    parser.add_argument("-b", "--noOfBins", dest="noOfBins", type=int, required=False,
                        help="OPTIONAL: number of equal-occupancy (count) bins. Default is 100")
    parser.add_argument("-m", "--mappabilityThres", dest="mappabilityThreshold", type=int, required=False,
                        help="OPTIONAL: minimum number of hits per locus that has to exist to call it mappable. "
                             "DEFAULT is 1.")
    parser.add_argument("-l", "--lib", dest="libname", required=False,
                        help="OPTIONAL: Name of the library that is analyzed to be used for name of file prefixes. "
                             "DEFAULT is FitHiC")
    parser.add_argument("-U", "--upperbound", dest="distUpThres", type=int, required=False,
                        help="OPTIONAL: upper bound on the intra-chromosomal distance range (unit: base pairs). "
                             "DEFAULT no limit.")
    parser.add_argument("-L", "--lowerbound", dest="distLowThres", type=int, required=False,
                        help="OPTIONAL: lower bound on the intra-chromosomal distance range (unit: base pairs). "
                             "DEFAULT no limit.")
    parser.add_argument("-v", "--visual", action="store_true", dest="visual", required=False,
                        help="OPTIONAL: plots (not produced by this implementation)")
    parser.add_argument("-x", "--contactType", dest="contactType", required=False,
                        help="OPTIONAL: which chromosomal regions to study (intraOnly, interOnly, All). "
                             "DEFAULT is intraOnly")
    parser.add_argument("-tL", "--biasLowerBound", dest="biasLowerBound", type=float, required=False,
                        help="OPTIONAL: lower bound of bias values to discard. DEFAULT is 0.5")
    parser.add_argument("-tU", "--biasUpperBound", dest="biasUpperBound", type=float, required=False,
                        help="OPTIONAL: upper bound of bias values to discard. DEFAULT is 2")
    parser.add_argument("-V", "--version", action="version", version="Fit-Hi-C (fithic_b200) {}".format(__version__),
                        help="Print version and exit")
    return parser.parse_args(args)


def _is_gz(path):
    """The reference probes the content (gzip.open(path).readline(), fithic/fithic.py:139-141), not the name."""
    try:
        with open(path, "rb") as f:
            return f.read(2) == b"\x1f\x8b"
    except OSError:
        return False


def settings_from_args(args):
    """Validation and defaults of main() (fithic/fithic.py:136-263); exits with status 2 like the reference."""
    print("\n")
    print("GIVEN FIT-HI-C ARGUMENTS")
    print("=========================")
    for label, path in (("interactions", args.intersfile), ("fragments", args.fragsfile)):
        if not os.path.exists(path):
            print("%s file not found" % label.capitalize())
            sys.exit(2)
        if not _is_gz(path):
            print("%s file must be gzipped (.gz)" % label.capitalize())
            sys.exit(2)
        print("Reading %s file from: %s" % (label, path))
    if not os.path.isdir(args.outdir):
        os.makedirs(args.outdir)
    print("Output path being used from %s" % args.outdir)
    if args.resolution < 0:
        print("Resolution must be a positive integer")
        sys.exit(2)
    if args.resolution == 0:
        print("Non-fixed size data being used.")  # restriction fragments (fithic/fithic.py:166-170)
    else:
        print("Fixed size data being used with resolution: %s" % args.resolution)
    if args.biasfile:
        if not os.path.exists(args.biasfile):
            print("Bias file not found")
            sys.exit(2)
        if not _is_gz(args.biasfile):
            print("Bias file must be gzipped (.gz)")
            sys.exit(2)
        print("Reading bias file from: %s" % args.biasfile)
    else:
        print("No bias file")
    st = Settings(resolution=args.resolution)
    # the reference's falsy-zero idiom: 0 means "use the default" (fithic/fithic.py:194-220)
    if args.noOfPasses:
        st.noOfPasses = args.noOfPasses
    print("The number of spline passes is %s" % st.noOfPasses)
    if args.noOfBins:
        st.noOfBins = args.noOfBins
    print("The number of bins is %s" % st.noOfBins)
    if args.mappabilityThreshold:
        st.mappThres = args.mappabilityThreshold
    print("The number of reads required to consider an interaction is %s" % st.mappThres)
    libName = args.libname if args.libname else "FitHiC"
    print("The name of the library for outputted files will be %s" % libName)
    if args.distUpThres:
        st.distUpThres = args.distUpThres
    if args.distLowThres:
        st.distLowThres = args.distLowThres
    print("Upper Distance threshold is %s" % st.distUpThres)
    print("Lower Distance threshold is %s" % st.distLowThres)
    if args.visual:
        print("Graphs are not produced by the B200 path (-v ignored)")
    region = args.contactType if args.contactType is not None else "intraOnly"
    if region == "All":
        print("All genomic regions will be analyzed")
        st.allReg = True
    elif region == "interOnly":
        print("Only inter-chromosomal regions will be analyzed")
        st.interOnly = True
    elif region == "intraOnly":
        print("Only intra-chromosomal regions will be analyzed")
    else:
        print("Invalid Option. Only options are 'All', 'interOnly', or 'intraOnly'")
        sys.exit(2)
    if args.biasLowerBound:
        st.biasLowerBound = args.biasLowerBound
    if args.biasUpperBound:
        st.biasUpperBound = args.biasUpperBound
    if st.biasLowerBound > st.biasUpperBound:
        print("Invalid Option. Bias lower bound is greater than bias upper bound. Please fix.")
        sys.exit(2)
    print("Lower bound of bias values is %s" % st.biasLowerBound)
    print("Upper bound of bias values is %s" % st.biasUpperBound)
    print("All arguments processed. Running FitHiC now...")
    print("=========================")
    print("\n")
    return st, libName


def shard_lines(chr_runs, n, world, snap=0.05):
    """Cut n file lines into `world` contiguous slices of about n / world lines; a cut moves to a chromosome boundary when
    one lies within snap * n / world lines of the even position.  Returns world + 1 cut positions."""
    bounds = np.cumsum(np.asarray(chr_runs[1], dtype=np.int64)) if chr_runs is not None and len(chr_runs[1]) else np.zeros(0)
    cuts = [0]
    for r in range(1, world):
        ideal = (n * r) // world
        if len(bounds):
            j = int(np.argmin(np.abs(bounds - ideal)))
            if abs(int(bounds[j]) - ideal) <= snap * n / world:
                ideal = int(bounds[j])
        cuts.append(max(cuts[-1], min(ideal, n)))
    cuts.append(n)
    return cuts


def run(contacts_path, frags_path, outdir, st, libName, bias_path=None, quiet=False):
    """Everything main() does after argument parsing.  Returns the per-pass result dicts (host numpy p/q/expcc of this
    rank's lines; under torchrun every rank runs this and rank 0 writes the tables and joins the output)."""
    import torch
    from .engine import Contacts, chr_runs_of
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dctx = None
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from .parallel import DistCtx
        dctx = DistCtx(torch.device("cuda", local))
    say = (lambda *a: None) if (quiet or rank != 0) else print
    t_run = time.time()
    t0 = time.time()
    say("Reading the contact counts file to generate bins...")
    contacts = fio.read_contacts(contacts_path)
    chroms = list(contacts.chroms)
    say("Interactions file read. Time took %s" % (time.time() - t0))
    t_read = time.time() - t0
    t1 = time.time()
    frags = fio.read_fragments(frags_path, chroms, st.mappThres, keep_mids=(st.resolution == 0))
    say("Fragments file read. Time took %s" % (time.time() - t1))
    biases, bias_log = None, []
    if bias_path:
        t1 = time.time()
        biases, bias_log = fio.read_biases(bias_path, chroms, st.resolution, st.biasLowerBound, st.biasUpperBound)
        say("Bias file read. Time took %s" % (time.time() - t1))
    contacts.chroms = chroms
    logfile = os.path.join(outdir, libName + ".fithic.log")
    n_file = len(contacts)
    lo, hi = 0, n_file
    if world > 1:  # this rank's slice of the file (every rank parsed the whole file: parsing is not the sharded part)
        cuts = shard_lines(contacts.chr_runs, n_file, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        sl = slice(lo, hi)
        contacts = Contacts(contacts.mid1[sl], contacts.mid2[sl], contacts.cnt[sl], contacts.chrs[sl], chroms,
                            chr_runs_of(contacts.chrs[sl]))

    eng = Engine(st, frags, biases, dist_ctx=dctx)
    eng.upload_contacts(contacts)
    eng.set_line_runs([lo], [hi - lo])
    outl, stats = eng.new_outlier_state()
    results = []
    metrics = {"world_size": world, "lines": n_file, "read_contacts_s": t_read, "passes": []}
    for passNo in range(1, st.noOfPasses + 1):
        if passNo > 1 and st.interOnly:
            say("Extra spline fits will not help with interOnly spline fit... Bypassing option")
            break
        ts = time.time()
        say("Spline fit Pass %s starting..." % passNo)
        r = eng.run_pass(passNo, outl, stats)
        torch.cuda.synchronize()
        t_gpu = time.time() - ts
        p = r["p"].cpu().numpy()
        q = r["q"].cpu().numpy()
        e = r["expcc"].cpu().numpy()
        n_out = int(stats[0].item())
        if dctx is not None:
            n_out = int(dctx.allreduce_small(np.array([n_out]))[0])
        r.update(p=p, q=q, expcc=e, n_outliers_total=n_out)
        say("Outlier threshold is... %s" % r["outlierThres"])
        suffix = ".res" + str(st.resolution) if st.resolution else ""  # -r 0 omits the part (:851, :1171)
        tab = os.path.join(outdir, libName + ".fithic_pass" + str(passNo) + suffix + ".txt")
        if rank == 0:
            write_log(logfile, r, st, frags, bias_log if passNo == 1 else [], tab)
            # bin table
            say("Writing %s" % tab)
            with open(tab, "w") as out:
                out.write("avgGenomicDist\tcontactProbability\tstandardError\tnoOfLocusPairs\ttotalOfContactCounts\n")
                b = r["bins"]
                for i in range(b["n"]):
                    out.write("%d\t%.2e\t%.2e\t%d\t%d\n" % (r["x_bins"][i], r["y_bins"][i], 0, b["pairs"][i], b["sumcc"][i]))
        sig = os.path.join(outdir, libName + ".spline_pass" + str(passNo) + suffix + ".significances.txt.gz")
        say("Writing p-values and q-values to file %s" % sig[:-3])
        # gzip level 2 by default: deflate, not formatting, bounds the writer (0.5 M rows/s/thread at level 2, 0.12 M at
        # level 6; the reference's level 9 manages 0.04 M rows/s); FITHIC_GZIP_LEVEL overrides
        tw = time.time()
        level = int(os.environ.get("FITHIC_GZIP_LEVEL", "2"))
        if world == 1:
            rows = fio.write_significances_native(sig, contacts, p, q, e, biases, st, level=level)
        else:
            import torch.distributed as dist
            part = "%s.part%04d" % (sig, rank)
            rows = fio.write_significances_native(part, contacts, p, q, e, biases, st, level=level, header=(rank == 0))
            rows = int(dctx.allreduce_small(np.array([rows]))[0])
            dist.barrier()
            if rank == 0:  # concatenated gzip members are one gzip file; the slices are in file order
                with open(sig, "wb") as out:
                    for k in range(world):
                        pk = "%s.part%04d" % (sig, k)
                        with open(pk, "rb") as f:
                            while True:
                                blk = f.read(1 << 24)
                                if not blk:
                                    break
                                out.write(blk)
                        os.remove(pk)
            dist.barrier()
        t_write = time.time() - tw
        say("Number of outliers is... %s" % r["n_outliers_total"])
        say("Spline fit Pass %s completed. Time took %s" % (passNo, time.time() - ts))
        metrics["passes"].append({"pass": passNo, "N": int(r["N"]), "T": r["T"], "bins": int(r["bins"]["n"]),
                                  "outlier_threshold": r["outlierThres"], "outliers_total": int(n_out), "rows_written": int(rows),
                                  "gpu_pass_s": t_gpu, "write_s": t_write,
                                  "host_ms": {k: v * 1e3 for k, v in eng.timings.get(passNo, {}).items()}})
        results.append(r)
    metrics["wall_s"] = time.time() - t_run
    if rank == 0:  # machine-readable run summary next to the log (the reference only has the text log)
        import json
        with open(os.path.join(outdir, libName + ".fithic_metrics.json"), "w") as f:
            json.dump(metrics, f, indent=1, default=float)
    say("=========================")
    say("Fit-Hi-C completed successfully")
    say("\n")
    return results


def main(argv=None):
    args = parse_args(sys.argv[1:] if argv is None else argv)
    if int(os.environ.get("RANK", "0")) != 0:  # under torchrun only rank 0 talks
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            st, libName = settings_from_args(args)
    else:
        st, libName = settings_from_args(args)
    run(args.intersfile, args.fragsfile, args.outdir, st, libName, args.biasfile)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
