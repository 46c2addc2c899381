"""The reference's function-level interface on the B200 path.

Same names, arguments and return shapes as fithic/fithic.py: read_Interactions,
makeBinsFromInteractions, generate_FragPairs, read_biases, calculateProbabilities, fit_Spline -- each one replaces the
Python loop of its namesake by the corresponding C-ABI call, so a driver written against the reference (its main(), or
a test harness that calls the stages one by one) runs unchanged.  Like the reference, the functions read module-level
settings (distLowThres, distUpThres, mappThres, interOnly, allReg, biasLowerBound, biasUpperBound, logfile); set them
on this module the way main() sets its globals.  Contacts are parsed and copied to the GPU once per file and stay
there across calls and passes.
"""
import numpy as np
import torch

from . import _capi
from . import io as fio
from .engine import (Engine, Fragments, Settings, calculate_probabilities, fit_spline, frag_pairs, make_bins)

# ---- module globals, as in the reference (fithic/fithic.py:193-307) ----
distLowThres = 0
distUpThres = float("inf")
mappThres = 1
interOnly = False
allReg = False
biasLowerBound = 0.5
biasUpperBound = 2
logfile = None
noOfBins = 100

_session = {}  # per contact file: parsed contacts, fragments, engine, outlier state


def _settings(resolution):
    return Settings(resolution=int(resolution), noOfBins=int(noOfBins), mappThres=int(mappThres), distLowThres=distLowThres,
                    distUpThres=distUpThres, interOnly=bool(interOnly), allReg=bool(allReg),
                    biasLowerBound=float(biasLowerBound), biasUpperBound=float(biasUpperBound))


def _contacts(path):
    s = _session.get(path)
    if s is None:
        s = _session[path] = dict(contacts=fio.read_contacts(path))
        s["chroms"] = list(s["contacts"].chroms)
    return s


def _log(text, mode="a"):
    if logfile:
        with open(logfile, mode) as f:
            f.write(text)


def _engine(s, resolution, frags=None, biases=None):
    """(Re)build the engine of a session when the resolution, the fragments or the biases become known."""
    key = (int(resolution), id(frags), id(biases))
    if s.get("engine_key") != key:
        fr = frags
        if fr is None:  # read_Interactions comes before generate_FragPairs: K1 does not look at the fragments
            fr = Fragments(list(s["chroms"]), np.zeros(len(s["chroms"]), np.int64), np.full(len(s["chroms"]), -1, np.int64))
            if not int(resolution):
                fr.mids = [np.zeros(0, np.int64) for _ in s["chroms"]]
        eng = Engine(_settings(resolution), fr, biases)
        eng.upload_contacts(s["contacts"])
        s["engine"], s["engine_key"] = eng, key
    s["engine"].st = _settings(resolution)
    return s["engine"]


def _outlier_state(eng, outliersline):
    """SortedList of line indices (duplicates allowed) -> per-line multiplicity + the reference's stalled-pointer limit
    (fithic/fithic.py:408-412)."""
    n = eng.n
    lines = np.asarray(list(outliersline), dtype=np.int64)
    mult = np.bincount(lines, minlength=n).astype(np.uint8) if len(lines) else np.zeros(n, np.uint8)
    dup = np.nonzero(mult >= 2)[0]
    limit = int(dup[0]) if len(dup) else n
    return torch.from_numpy(mult).to(eng.device), limit


def read_Interactions(contactCountsFile, biasFile, outliers=None):
    """fithic/fithic.py:389-454.  Returns (mainDic, observedInterAllCount, observedInterAllSum, observedIntraAllSum,
    observedIntraInRangeSum) with mainDic = {distance: [0, sum of counts]}."""
    s = _contacts(contactCountsFile)
    res = s.get("resolution")
    if res is None:
        raise RuntimeError("set the resolution first: refapi.set_resolution(contactCountsFile, resolution)")
    eng = _engine(s, res, s.get("frags"), s.get("biases"))
    skip, limit = (None, -1)
    if outliers is not None and len(outliers):
        skip, limit = _outlier_state(eng, outliers)
    hist, present, scal = eng.hist_distance(skip, limit)
    D = eng.D
    h = hist.cpu().numpy()
    bits = np.unpackbits(present.cpu().numpy().view(np.uint8), bitorder="little")[:D].astype(bool)
    sc = scal.cpu().numpy()
    if int(sc[_capi.S_OFFGRID]):
        raise ValueError("contact distances off the %d bp grid are not supported" % res)
    seen = np.nonzero((h != 0) | bits)[0]
    mainDic = {int(k) * eng.grid: [0, int(h[k])] for k in seen}
    _log("\n\nInteractions file read successfully\n" + "-" * 84 + "\n"
         "Observed, Intra-chr in range: pairs= %d\t totalCount= %d\n" % (sc[_capi.S_INTRA_INRANGE_LINES], sc[0]) +
         "Observed, Intra-chr all: pairs= %d\t totalCount= %d\n" % (sc[_capi.S_INTRA_ALL_LINES], sc[1]) +
         "Observed, Inter-chr all: pairs= %d\t totalCount= %d\n\n" % (sc[3], sc[2]), "w")
    s["max_count"] = int(sc[_capi.S_MAX_COUNT])
    return (mainDic, int(sc[_capi.S_INTER_ALL_COUNT]), int(sc[_capi.S_INTER_ALL_SUM]), int(sc[_capi.S_INTRA_ALL_SUM]),
            int(sc[_capi.S_INTRA_INRANGE_SUM]))


def set_resolution(contactCountsFile, resolution):
    """The reference learns the resolution only in generate_FragPairs; the dense histogram needs it from the start."""
    _contacts(contactCountsFile)["resolution"] = int(resolution)


def makeBinsFromInteractions(mainDic, noOfBins_, observedIntraInRangeSum, outliersdist=None):
    """fithic/fithic.py:463-553.  binStats[i] = [(lb, ub), pairs, sumCC, sumDist, avgCC, avgDist, [distances], pairs]."""
    lib = _capi.load()
    dists = np.array(sorted(mainDic.keys()), dtype=np.int64)
    sums = np.array([mainDic[int(d)][1] for d in dists], dtype=np.int64)
    b = make_bins(lib, dists, sums, int(noOfBins_), int(observedIntraInRangeSum))
    binStats = {}
    for i in range(b["n"]):
        inside = dists[(dists >= b["lb"][i] if i else dists >= 0) & (dists <= b["ub"][i]) &
                       (dists > (b["ub"][i - 1] if i else -1))]
        binStats[i] = [(int(b["lb"][i]), int(b["ub"][i])), 0, int(b["sumcc"][i]), 0, 0, 0, [int(d) for d in inside], 0]
    if outliersdist is not None and len(binStats):
        ub = np.array([binStats[i][0][1] for i in range(len(binStats))], dtype=np.int64)
        od = np.asarray(list(outliersdist), dtype=np.int64)
        which = np.minimum(np.searchsorted(ub, od, side="left"), len(ub) - 1)
        dec = np.bincount(which, minlength=len(ub))
        for i in range(len(binStats)):
            binStats[i][1] -= int(dec[i])
            binStats[i][7] -= int(dec[i])
    _log("Equal occupancy bins generated\n\n")
    return binStats


def generate_FragPairs(observedInterAllCount, observedInterAllSum, binStats, fragsfile, resolution):
    """fithic/fithic.py:561-793, both branches (fixed-size bins :596-689, restriction fragments with resolution 0
    :691-778).  Returns (binStats, noOfFrags, maxPossibleGenomicDist, possibleIntraInRangeCount, possibleInterAllCount,
    interChrProb, baselineIntraChrProb)."""
    lib = _capi.load()
    chroms = []
    for s in _session.values():
        chroms = s["chroms"]
        break
    frags = fio.read_fragments(fragsfile, chroms, mappThres, keep_mids=not resolution)
    for s in _session.values():
        s["frags"] = frags
    st = _settings(resolution)
    nb = len(binStats)
    bins = dict(n=nb, lb=np.array([binStats[i][0][0] for i in range(nb)], dtype=np.int64),
                ub=np.array([binStats[i][0][1] for i in range(nb)], dtype=np.int64),
                sumcc=np.array([binStats[i][2] for i in range(nb)], dtype=np.int64))
    dec = -np.array([binStats[i][1] for i in range(nb)], dtype=np.int64) if nb else None
    fp = frag_pairs(lib, frags, st, bins, dec)
    for i in range(nb):
        binStats[i][1] = int(bins["pairs"][i])
        binStats[i][7] = int(bins["pairs7"][i])
        binStats[i][3] = float(bins["sumdist"][i])
    ok = frags.n_mappable > 0
    if not resolution:
        maxd = float(fp["maxPossibleGenomicDist"])  # the largest in-range fragment distance (:712)
    else:
        maxd = float((frags.max_mid[ok] - resolution / 2).max()) if ok.any() else 0
    interChrProb = 1.0 / observedInterAllCount if observedInterAllCount > 0 else 0
    pia = fp["possibleIntraAllCount"]
    return (binStats, fp["noOfFrags"], maxd, fp["possibleIntraInRangeCount"], fp["possibleInterAllCount"], interChrProb,
            1.0 / pia if pia > 0 else 0)


def read_biases(infilename):
    """fithic/fithic.py:798-837.  Returns the dense per-locus table the kernels use (truthy, like the reference's dict)."""
    res = None
    chroms = []
    for s in _session.values():
        res, chroms = s.get("resolution"), s["chroms"]
        break
    b, log = fio.read_biases(infilename, chroms, res, float(biasLowerBound), float(biasUpperBound))
    for s in _session.values():
        s["biases"] = b
    _log("\n".join(log) + "\n\n")
    return b


def calculateProbabilities(mainDic, binStats, resolution, outfilename, observedIntraInRangeSum):
    """fithic/fithic.py:843-918.  Returns [x, y, yerr] and writes `${outfilename}.res${R}.txt`."""
    nb = len(binStats)
    bins = dict(n=nb, pairs=np.array([binStats[i][1] for i in range(nb)], dtype=np.int64),
                pairs7=np.array([binStats[i][7] for i in range(nb)], dtype=np.int64),
                sumcc=np.array([binStats[i][2] for i in range(nb)], dtype=np.int64),
                sumdist=np.array([binStats[i][3] for i in range(nb)], dtype=np.float64))
    x, y = calculate_probabilities(bins, observedIntraInRangeSum)
    for i in range(nb):
        binStats[i][4], binStats[i][5] = y[i], x[i]
    name = outfilename + (".res" + str(resolution) if resolution else "") + ".txt"
    with open(name, "w") as out:
        out.write("avgGenomicDist\tcontactProbability\tstandardError\tnoOfLocusPairs\ttotalOfContactCounts\n")
        for i in range(nb):
            out.write("%d\t%.2e\t%.2e\t%d\t%d\n" % (x[i], y[i], 0, binStats[i][1], binStats[i][2]))
    _log("Means and error written to %s\n\n" % name)
    return [x, y, [0] * nb]


def fit_Spline(mainDic, x, y, yerr, infilename, outfilename, biasDic, outliersline, outliersdist, observedIntraInRangeSum,
               possibleIntraInRangeCount, possibleInterAllCount, observedInterAllCount, observedIntraAllSum,
               observedInterAllSum, biasLowerBound_, biasUpperBound_, resolution, passNo):
    """fithic/fithic.py:925-1233.  Fits the spline (scipy, <= noOfBins points), evaluates it, scores every line of
    `infilename` (K2, K3), corrects (K4), writes `${outfilename}.res${R}.significances.txt.gz`, extends the outlier
    lists.  Returns [splineX, newSplineY, residual, outliersline, outliersdist, FDRx, FDRy] (the last two are plot data
    of the reference and come back empty)."""
    s = _contacts(infilename)
    biases = biasDic if biasDic else None
    eng = _engine(s, resolution, s.get("frags"), biases)
    st = eng.st
    dists = np.array(sorted(mainDic.keys()), dtype=np.int64)
    splineX, table, lut = None, None, None
    if not st.interOnly:
        xs, ys, tck = fit_spline(x, y)
        splineX = dists[(dists >= min(xs)) & (dists <= max(xs))]
        table, lut = eng.spline_table(tck, splineX, min(xs), max(xs))
    if st.allReg:
        T = possibleIntraInRangeCount + observedInterAllCount
    elif st.interOnly:
        T = observedInterAllCount
    else:
        T = possibleIntraInRangeCount
    thres = 1.0 / T
    print("Outlier threshold is... %s" % thres)
    outl = torch.zeros(max(eng.n, 1), dtype=torch.uint8, device=eng.device)[:eng.n]
    stats = torch.tensor([0, -1], dtype=torch.int64, device=eng.device)
    interChrProb = 1.0 / observedInterAllCount if observedInterAllCount > 0 else 0.0
    p, e = eng.pvalues(lut, observedIntraInRangeSum, observedInterAllSum, interChrProb, s.get("max_count", 1 << 20), outl,
                       thres, stats)
    q = eng.bh_qvalues(p, float(T))
    torch.cuda.synchronize()
    ph, qh, eh = p.cpu().numpy(), q.cpu().numpy(), e.cpu().numpy()
    sig = outfilename + (".res" + str(resolution) if resolution else "") + ".significances.txt.gz"
    print("Writing p-values and q-values to file %s" % (outfilename + ".significances.txt"))
    fio.write_significances_native(sig, s["contacts"], ph, qh, eh, biases, st)
    flagged = np.nonzero(outl.cpu().numpy())[0]
    c = s["contacts"]
    d = np.abs(c.mid1[flagged].astype(np.int64) - c.mid2[flagged].astype(np.int64))
    for i, dd in zip(flagged.tolist(), d.tolist()):
        outliersline.add(i)
        outliersdist.add(dd)
    _log("Spline successfully fit\n\n\n")
    s["last"] = dict(p=ph, q=qh, expcc=eh, T=T)
    newY = table.cpu().numpy().copy() if table is not None else None
    return [None if splineX is None else [int(v) for v in splineX], newY, 0, outliersline, outliersdist, [], []]


def last_results(contactCountsFile):
    """Full-precision p, q, ExpCC of the last fit_Spline call on this file (the output file keeps 7 digits)."""
    return _session[contactCountsFile]["last"]


def reset():
    _session.clear()
