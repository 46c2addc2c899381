"""Text boundary of the path: gzipped contact / fragment / bias files in, `.significances.txt.gz` out.

Formats follow the reference (fithic/fithic.py:413, :580-590, :807-832, :1166-1212).  Parsing and formatting are host
work (the reference spends most of its wall time here); the arrays they produce are what the kernels consume.
"""
import gzip
import math
import os

import numpy as np
import pandas as pd

from .engine import Biases, Contacts, Fragments, chr_runs_of

_I32_MAX = (1 << 31) - 1


def _open_text(path):
    return gzip.open(path, "rt")


def read_contacts(path):
    """contactCounts file: chr1 mid1 chr2 mid2 count, whitespace separated (fithic/fithic.py:413-417).
    Native reader (csrc/textio.cu): one thread inflates, one parses straight into the int32 arrays."""
    from . import _capi
    lib = _capi.load()
    h = lib.fhc_io_read_contacts(os.fsencode(path))
    if not h:
        raise ValueError(lib.fhc_last_error().decode("utf-8", "replace"))
    try:
        n = int(lib.fhc_io_contacts_n(h))
        chroms = [lib.fhc_io_contacts_chrom(h, i).decode() for i in range(int(lib.fhc_io_contacts_nchrom(h)))]
        m1 = np.empty(n, dtype=np.int32)
        m2 = np.empty(n, dtype=np.int32)
        cnt = np.empty(n, dtype=np.int32)
        chrs = np.empty(n, dtype=np.uint32)
        _capi.check(lib.fhc_io_contacts_copy(h, _capi.dptr(m1), _capi.dptr(m2), _capi.dptr(cnt), _capi.dptr(chrs)))
    finally:
        lib.fhc_io_free(h)
    return Contacts(m1, m2, cnt, chrs, chroms, chr_runs_of(chrs))


def read_contacts_pandas(path):
    """The same through pandas (kept to cross-check the native reader in tests)."""
    df = pd.read_csv(path, sep=r"\s+", header=None, names=["c1", "m1", "c2", "m2", "n"], engine="c",
                     dtype={"c1": str, "c2": str, "m1": np.int64, "m2": np.int64, "n": np.float64},
                     compression="gzip", float_precision="round_trip")
    chroms = list(pd.unique(pd.concat([df.c1, df.c2], ignore_index=True)))
    if len(chroms) >= (1 << 16):
        raise ValueError("more than 65535 chromosome names")
    cid = {c: i for i, c in enumerate(chroms)}
    c1 = df.c1.map(cid).to_numpy(np.uint32)
    c2 = df.c2.map(cid).to_numpy(np.uint32)
    m1 = df.m1.to_numpy()
    m2 = df.m2.to_numpy()
    cnt = np.trunc(df.n.to_numpy())  # int(float(text)): truncation toward zero
    if len(m1) and (max(m1.max(), m2.max()) > _I32_MAX or min(m1.min(), m2.min()) < 0 or np.abs(cnt).max() > _I32_MAX):
        raise ValueError("mid points / counts outside int32")
    return Contacts(m1.astype(np.int32), m2.astype(np.int32), cnt.astype(np.int32), (c1 | (c2 << 16)).astype(np.uint32),
                    chroms)


def read_fragments(path, chroms, mappThres, keep_mids=False):
    """fragments file: uses column 0 (chr), 2 (mid) and 3 (hit count >= mappThres means mappable; :583-590).
    `chroms` (list) is extended in place by chromosomes that only occur here.  keep_mids: also keep every chromosome's
    sorted mappable mid points (restriction-fragment mode, -r 0, enumerates fragment pairs, :691-778)."""
    df = pd.read_csv(path, sep=r"\s+", header=None, engine="c", usecols=[0, 2, 3], names=["c", "m", "h"],
                     dtype={"c": str, "m": np.int64, "h": np.int64}, compression="gzip")
    cid = {c: i for i, c in enumerate(chroms)}
    for c in pd.unique(df.c):
        if c not in cid:
            cid[c] = len(chroms)
            chroms.append(c)
    ok = df[df.h >= mappThres]
    ids = ok.c.map(cid).to_numpy(np.int64)
    n = np.bincount(ids, minlength=len(chroms)).astype(np.int64)
    mx = np.full(len(chroms), -1, dtype=np.int64)
    np.maximum.at(mx, ids, ok.m.to_numpy(np.int64))
    mids = None
    if keep_mids:
        allm = ok.m.to_numpy(np.int64)
        mids = [np.sort(allm[ids == ci]) for ci in range(len(chroms))]
    return Fragments(list(chroms), n, mx, mids)


def _mquantiles(a, prob, alphap=0.4, betap=0.4):
    """scipy.stats.mstats.mquantiles(a, prob) for a 1-d array without masked entries (what read_biases logs,
    fithic/fithic.py:812-816), restated so that the command line does not pay for importing scipy.stats (0.5 s): the
    plotting positions of Hyndman & Fan's family, (1 - g) x[k-1] + g x[k] on the sorted data with
    aleph = n p + alphap + p (1 - alphap - betap), k = floor(clip(aleph, 1, n - 1)), g = clip(aleph - k, 0, 1)."""
    x = np.sort(np.asarray(a, dtype=np.float64))
    n = len(x)
    p = np.asarray(prob, dtype=np.float64)
    if n == 1:  # scipy's special case: the value itself, no interpolation arithmetic
        return np.resize(x, p.shape)
    m = alphap + p * (1.0 - alphap - betap)
    aleph = n * p + m
    k = np.floor(aleph.clip(1, n - 1)).astype(int)
    gamma = (aleph - k).clip(0, 1)
    return (1.0 - gamma) * x[(k - 1).tolist()] + gamma * x[k.tolist()]


def read_biases(path, chroms, resolution, biasLowerBound, biasUpperBound):
    """bias file: chr mid bias (fithic/fithic.py:798-837).  bias < tL, NaN or > tU -> -1; FIRST occurrence of a
    (chr, mid) wins.  Returns (Biases, log lines); `chroms` is extended in place."""
    df = pd.read_csv(path, sep=r"\s+", header=None, engine="c", names=["c", "m", "b"],
                     dtype={"c": str, "m": np.int64, "b": np.float64}, compression="gzip",
                     float_precision="round_trip")  # Python float() semantics, like the reference
    cid = {c: i for i, c in enumerate(chroms)}
    for c in pd.unique(df.c):
        if c not in cid:
            cid[c] = len(chroms)
            chroms.append(c)
    ids = df.c.map(cid).to_numpy(np.int64)
    mids = df.m.to_numpy(np.int64)
    b = df.b.to_numpy(np.float64).copy()
    raw = b[b != 1.0]
    log = []
    if len(raw):
        botQ, med, topQ = _mquantiles(raw, (0.05, 0.5, 0.95))
        log += ["5th quantile of biases: %s" % botQ, "50th quantile of biases: %s" % med,
                "95th quantile of biases: %s" % topQ]
    bad = (b < biasLowerBound) | np.isnan(b) | (b > biasUpperBound)
    b[bad] = -1.0
    log.append("Out of %d loci %d were discarded with biases not in range [%s-%s]" %
               (len(b), int(bad.sum()), biasLowerBound, biasUpperBound))
    if len(mids) and (mids.min() < 0 or mids.max() > _I32_MAX):
        raise ValueError("bias mid points outside int32")
    nchr = len(chroms)
    if resolution == 0:
        # restriction fragments: no grid.  Per chromosome the loci in ascending mid order, first occurrence of a repeated
        # (chr, mid) wins (:823-829); K3 and the writer find a locus by binary search.
        order = np.lexsort((np.arange(len(mids)), mids, ids))  # by chromosome, mid, then file order
        ids_s, mids_s, b_s = ids[order], mids[order], b[order]
        first = np.ones(len(order), dtype=bool)
        first[1:] = (ids_s[1:] != ids_s[:-1]) | (mids_s[1:] != mids_s[:-1])
        ids_s, mids_s, b_s = ids_s[first], mids_s[first], b_s[first]
        chr_off = np.zeros(nchr + 1, dtype=np.int64)
        np.cumsum(np.bincount(ids_s, minlength=nchr), out=chr_off[1:])
        return Biases(np.ascontiguousarray(b_s, dtype=np.float64), np.ascontiguousarray(mids_s, dtype=np.int32), chr_off,
                      True), log
    nslot = np.zeros(nchr, dtype=np.int64)
    if len(mids):
        np.maximum.at(nslot, ids, mids // resolution + 1)
    chr_off = np.zeros(nchr + 1, dtype=np.int64)
    np.cumsum(nslot, out=chr_off[1:])
    slot = chr_off[ids] + mids // resolution
    values = np.full(int(chr_off[-1]), -1.0, dtype=np.float64)
    smid = np.full(int(chr_off[-1]), -1, dtype=np.int32)
    # first occurrence wins: write in reverse so that the earliest line lands last
    values[slot[::-1]] = b[::-1]
    smid[slot[::-1]] = mids[::-1].astype(np.int32)
    # two different mid points in one slot cannot be represented by the dense table
    if np.any(smid[slot] != mids):
        raise ValueError("bias file has several mid points inside one %d bp bin; only one locus per bin is supported"
                         % resolution)
    return Biases(values, smid, chr_off), log


def lookup_biases(biases, chr_ids, mids, resolution):
    """Host gather of the per-line bias column of the output (missing locus -> -1, :1026-1054)."""
    if biases is None:
        return np.ones(len(mids), dtype=np.float64)
    nchr = len(biases.chr_off) - 1
    chr_ids = chr_ids.astype(np.int64)
    ok = chr_ids < nchr
    cidc = np.where(ok, chr_ids, 0)
    if biases.sparse:
        out = np.full(len(mids), -1.0)
        for c in np.unique(cidc[ok]):
            lo, hi = int(biases.chr_off[c]), int(biases.chr_off[c + 1])
            sel = np.nonzero(ok & (cidc == c))[0]
            if hi > lo and len(sel):
                pos = np.minimum(np.searchsorted(biases.mids[lo:hi], mids[sel]), hi - lo - 1)
                hit = biases.mids[lo:hi][pos] == mids[sel]
                out[sel[hit]] = biases.values[lo:hi][pos[hit]]
        return out
    slot = biases.chr_off[cidc] + mids.astype(np.int64) // resolution
    ok &= slot < biases.chr_off[cidc + 1]
    slot = np.where(ok, slot, 0)
    if len(biases.mids):
        ok &= biases.mids[slot] == mids
        return np.where(ok, biases.values[slot], -1.0)
    return np.full(len(mids), -1.0)


def write_significances_native(path, contacts, p, q, expcc, biases, settings, nthreads=None, level=6, header=True):
    """`.significances.txt.gz` (fithic/fithic.py:1166-1212) through the native multi-threaded formatter + gzip
    (csrc/textio.cu).  Returns the number of rows written."""
    import ctypes
    from . import _capi
    lib = _capi.load()
    st = settings
    n = len(contacts)
    names = (ctypes.c_char_p * len(contacts.chroms))(*[c.encode() for c in contacts.chroms])
    arr = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    m1, m2, cnt, chrs = arr(contacts.mid1, np.int32), arr(contacts.mid2, np.int32), arr(contacts.cnt, np.int32), \
        arr(contacts.chrs, np.uint32)
    p, q, e = arr(p, np.float64), arr(q, np.float64), arr(expcc, np.float64)
    mode = _capi.MODE_ALL if st.allReg else (_capi.MODE_INTER_ONLY if st.interOnly else _capi.MODE_INTRA_ONLY)
    U = -1 if math.isinf(st.distUpThres) else int(st.distUpThres)
    bv = bm = bo = None
    nb = 0
    if biases is not None:
        bv, bm, bo = arr(biases.values, np.float64), arr(biases.mids, np.int32), arr(biases.chr_off, np.int64)
        nb = len(bo) - 1
    if nthreads is None:
        nthreads = min(os.cpu_count() or 1, 32)
    rows = lib.fhc_io_write_significances(os.fsencode(path), names, len(contacts.chroms), _capi.dptr(m1), _capi.dptr(m2),
                                          _capi.dptr(cnt), _capi.dptr(chrs), _capi.dptr(p), _capi.dptr(q), _capi.dptr(e),
                                          n, mode, int(st.distLowThres), U, _capi.dptr(bv), _capi.dptr(bm), _capi.dptr(bo),
                                          nb, int(st.resolution), int(nthreads), int(level), 1 if header else 0)
    return _capi.check(rows)


def write_significances(path, contacts, p, q, expcc, bias1, bias2, settings, chunk=1 << 18):
    """The same in pure Python (the reference's own formatting expression; used to cross-check the native writer)."""
    st = settings
    c1 = (contacts.chrs & 0xffff).astype(np.int64)
    c2 = (contacts.chrs >> 16).astype(np.int64)
    intra = c1 == c2
    d = np.abs(contacts.mid1.astype(np.int64) - contacts.mid2.astype(np.int64))
    L, U = st.distLowThres, st.distUpThres
    inrange = ((L == -1) | ((L > -1) & (d >= L))) & ((U == -1) | ((U > -1) & (d <= U)))
    keep = np.zeros(len(d), dtype=bool)
    if st.allReg or st.interOnly:
        keep |= ~intra                                    # :1197
    if st.allReg or not st.interOnly:
        keep |= intra & inrange                           # :1205-1207
    idx = np.nonzero(keep)[0]
    names = np.asarray(contacts.chroms, dtype=object)
    with gzip.open(path, "wt") as out:
        out.write("chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\tbias1\tbias2\tExpCC\n")
        for s in range(0, len(idx), chunk):
            ii = idx[s:s + chunk]
            rows = zip(names[c1[ii]], contacts.mid1[ii].tolist(), names[c2[ii]], contacts.mid2[ii].tolist(),
                       contacts.cnt[ii].tolist(), p[ii].tolist(), q[ii].tolist(), bias1[ii].tolist(),
                       bias2[ii].tolist(), expcc[ii].tolist())
            out.write("".join("%s\t%d\t%s\t%d\t%d\t%e\t%e\t%e\t%e\t%f\n" % r for r in rows))
    return len(idx)
