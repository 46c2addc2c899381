"""Device-resident significance engine: K1 -> host binning + spline fit -> K2 -> K3 -> K4 for every spline pass.

This is the array-level host side of the hot path (reference fithic/fithic.py main() pass loop :317-376).  Contacts
live in HBM as four int32 arrays for the whole run; per pass the host sees only the distance histogram (<= 400 kB) and
sends back the spline knots.  All arithmetic on contacts happens in libfithic_b200.so; torch is used for device
memory, streams and (multi-GPU) torch.distributed collectives only.
"""
import ctypes
import math
import os
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _capi
from ._capi import check, dptr

_U64_MAX = (1 << 64) - 1


@dataclass
class Settings:
    """The module globals main() sets in the reference (fithic/fithic.py:193-260)."""
    resolution: int = 0
    noOfBins: int = 100
    mappThres: int = 1
    distLowThres: int = 0          # -1 = unbounded
    distUpThres: float = float("inf")
    interOnly: bool = False
    allReg: bool = False
    biasLowerBound: float = 0.5
    biasUpperBound: float = 2.0
    noOfPasses: int = 1

    @property
    def mode(self):
        if self.allReg:
            return _capi.MODE_ALL
        return _capi.MODE_INTER_ONLY if self.interOnly else _capi.MODE_INTRA_ONLY

    @property
    def L(self):
        return int(self.distLowThres)

    @property
    def U(self):
        return -1 if math.isinf(self.distUpThres) else int(self.distUpThres)


@dataclass
class Contacts:
    """Host structure of arrays, one element per line of the contact-counts file (file order)."""
    mid1: np.ndarray   # int32
    mid2: np.ndarray   # int32
    cnt: np.ndarray    # int32, int(float(text)) (fithic/fithic.py:415, myUtils.py:123-124)
    chrs: np.ndarray   # uint32, chr1 | chr2 << 16 (ids into `chroms`)
    chroms: list = field(default_factory=list)
    # Optional run-length form of `chrs`: (values uint32 [r], lengths int64 [r]).  Contact files are grouped by
    # chromosome, so this is a few dozen numbers; when it is present the 4 bytes per line of `chrs` never cross the PCIe
    # link (Engine.upload_contacts expands the runs on the device).  io.read_contacts and chr_runs_of() produce it.
    chr_runs: tuple = None

    def __len__(self):
        return int(self.mid1.shape[0])


def chr_runs_of(chrs):
    """Run-length encoding of a chrs array: (values uint32, lengths int64)."""
    chrs = np.asarray(chrs)
    if len(chrs) == 0:
        return np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.int64)
    cut = np.flatnonzero(chrs[1:] != chrs[:-1]) + 1
    starts = np.concatenate([[0], cut])
    return chrs[starts].astype(np.uint32), np.diff(np.concatenate([starts, [len(chrs)]])).astype(np.int64)


@dataclass
class Fragments:
    """What generate_FragPairs needs from the fragments file (fixed-size branch, fithic/fithic.py:580-604):
    per chromosome the number of mappable loci and their largest mid point.  Order = `chroms` order of Contacts,
    extended by chromosomes that only occur in the fragments file."""
    chroms: list
    n_mappable: np.ndarray  # int64 per chromosome
    max_mid: np.ndarray     # int64 per chromosome
    mids: list = None       # -r 0 only: per chromosome, the ascending mid points of its mappable fragments (int64 arrays)


@dataclass
class Biases:
    """Dense per-locus bias vector: slot(chr, mid) = chr_off[chr] + mid // res; -1 = discarded (read_biases, :818-832)."""
    values: np.ndarray   # float64 [nslots]
    mids: np.ndarray     # int32 [nslots], the mid point stored in the slot (-1 = empty)
    chr_off: np.ndarray  # int64 [nchr + 1]
    sparse: bool = False  # restriction-fragment mode (-r 0): no grid -- per chromosome the loci of the bias file in ascending
                          # mid order; a locus is found by binary search


class Engine:
    """Runs spline passes on one GPU.  With `dist_ctx` set (parallel.DistCtx), histograms/totals are all-reduced so that
    every rank fits the same spline, and q-values come from the range-partitioned global BH."""

    def __init__(self, settings, fragments, biases=None, device=None, dist_ctx=None):
        self.lib = _capi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("fithic_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.st = settings
        # Restriction-fragment mode (-r 0): distances are arbitrary integers.  The kernels work on a grid of `grid` bp; with
        # grid = 1 every distance is its own slot, so K1 (dense histogram), K2 (lookup table by slot) and K3 run unchanged --
        # the slot arrays just get as long as the largest in-range distance (5 M entries for -U 5000000, 40 MB each).
        self.grid = settings.resolution if settings.resolution > 0 else 1
        if settings.resolution == 0:
            if fragments.mids is None:
                raise ValueError("restriction-fragment mode (-r 0) needs the fragment mid points (io.read_fragments(..., keep_mids=True))")
            if biases is not None and not biases.sparse:
                raise ValueError("restriction-fragment mode (-r 0) needs the bias table in its sparse layout "
                                 "(io.read_biases(..., resolution=0))")
        self.frags = fragments
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dist = dist_ctx
        self.timings = {}
        self._bias_host = biases
        self._bias_dev = None
        if biases is not None:
            # fixed-size bins on the regular grid (every slot holds mid = k * res + res / 2 or nothing): K3 can check the
            # mid point arithmetically and skips the gather of the stored mid points
            regular = False
            if settings.resolution > 0 and len(biases.mids) and not biases.sparse:
                nslot = np.diff(biases.chr_off)
                k = np.arange(len(biases.mids), dtype=np.int64) - np.repeat(biases.chr_off[:-1], nslot)
                want = k * settings.resolution + settings.resolution // 2
                empty = biases.mids < 0
                regular = bool(np.all(empty | (biases.mids == want)) and np.all(biases.values[empty] == -1.0))
            self._bias_dev = (torch.from_numpy(biases.values).to(self.device),
                              None if regular else torch.from_numpy(biases.mids).to(self.device),
                              torch.from_numpy(biases.chr_off).to(self.device))
            self._bias_sparse = 1 if biases.sparse else 0
        self._ws = {}
        self.contacts = None
        self.n = 0

    # ------------------------------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _buf(self, name, nbytes):
        """Grow-only byte workspace."""
        cur = self._ws.get(name)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[name] = cur
        return cur

    def _tensor(self, name, n, dtype):
        cur = self._ws.get(name)
        if cur is None or cur.numel() < n or cur.dtype != dtype:
            cur = torch.empty(int(max(n, 1)), dtype=dtype, device=self.device)
            self._ws[name] = cur
        return cur[:n]

    # ------------------------------------------------------------------------------------------------------------
    def set_contacts_device(self, mid1, mid2, cnt, chrs):
        """Adopt device-resident int32 tensors (chrs holds the uint32 bit pattern)."""
        n = mid1.numel()
        for t in (mid1, mid2, cnt, chrs):
            assert t.is_cuda and t.dtype == torch.int32 and t.numel() == n and t.is_contiguous()
        self.contacts = (mid1, mid2, cnt, chrs)
        self.n = n
        self.D = None
        self._own_contacts = False

    def upload_contacts(self, c, non_blocking=False):
        """Host SoA -> HBM (the only per-contact host->device traffic of a run: 16 B per contact)."""
        ts = []
        reuse = self.contacts if (self.contacts is not None and getattr(self, "_own_contacts", False)
                                  and self.n == len(c)) else None
        runs = c.chr_runs if (c.chr_runs is not None and len(c.chr_runs[0]) <= self.MAX_CHR_RUNS) else None
        for j, a in enumerate((c.mid1, c.mid2, c.cnt, c.chrs.view(np.int32) if runs is None else None)):
            if a is None:
                # chromosome ids from their run-length form: nothing over the link instead of 4 B per line; one fill per run
                # (torch.repeat_interleave parallelises over the runs, not over the output: 30 ms for 24 runs of 12 M)
                assert int(np.sum(runs[1])) == len(c), "chr_runs do not cover the contacts"
                d = reuse[j] if reuse is not None else torch.empty(len(c), dtype=torch.int32, device=self.device)
                pos = 0
                for v, ln in zip(np.asarray(runs[0]).view(np.int32).tolist(), np.asarray(runs[1]).tolist()):
                    d[pos:pos + ln].fill_(v)
                    pos += ln
                ts.append(d)
                continue
            h = torch.from_numpy(np.ascontiguousarray(a))
            if reuse is not None:  # same size as the previous upload: no new device allocation
                reuse[j].copy_(h, non_blocking=non_blocking)
                ts.append(reuse[j])
            else:
                ts.append(h.to(self.device, non_blocking=non_blocking))
        self.set_contacts_device(*ts)
        self._own_contacts = True

    def distance_slots(self):
        """D = number of distance slots: every |mid1 - mid2| of an intra line is < D * res."""
        if self.D is None:
            mid1, mid2, _, _ = self.contacts
            if self.n == 0:
                mx = 0
            else:
                mx = int(torch.maximum(mid1.max(), mid2.max()).item()) - int(torch.minimum(mid1.min(), mid2.min()).item())
            if len(self.frags.max_mid):
                mx = max(mx, int(self.frags.max_mid.max()))
            if self.st.resolution == 0 and self.st.U >= 0:
                mx = min(mx, self.st.U)  # on the 1 bp grid only in-range distances need a slot
            self.D = mx // self.grid + 2
            if self.dist is not None:
                self.D = self.dist.max_int(self.D)  # the histogram is all-reduced: every rank needs the same length
        return self.D

    # ------------------------------------------------------------------------------------------------------------
    # K1  (read_Interactions, fithic/fithic.py:389-454)
    def hist_distance(self, skip=None, skip_limit=-1):
        D = self.distance_slots()
        mid1, mid2, cnt, chrs = self.contacts
        # one buffer [hist | totals | seen bitmap] so that the host needs a single device->host copy per pass
        nwords = (D + 31) // 32
        buf = self._tensor("k1buf", D + _capi.N_SCALARS + (nwords + 1) // 2, torch.int64)
        hist = buf[:D]
        scal = buf[D:D + _capi.N_SCALARS]
        present = buf[D + _capi.N_SCALARS:].view(torch.int32)[:nwords]
        check(self.lib.fhc_hist_distance(dptr(mid1), dptr(mid2), dptr(cnt), dptr(chrs), dptr(skip), int(skip_limit),
                                         self.n, self.st.L, self.st.U, self.grid, dptr(hist), dptr(present), D,
                                         dptr(scal), self._stream()))
        return hist, present, scal

    # ------------------------------------------------------------------------------------------------------------
    def run_pass(self, passNo, outl=None, outl_stats=None, after_pvalues=None, pvalue_chunks=1, after_chunk=None):
        """One spline pass.  Returns a dict with host-side tables and device tensors p, q, expcc."""
        st, lib = self.st, self.lib
        res = self.grid
        ev = {}
        t0 = time.perf_counter()
        # ---- K1 ----
        skip, skip_limit = None, -1
        if passNo > 1 and outl is not None:
            skip = outl
            first_dup = int(outl_stats[1].item()) & _U64_MAX
            skip_limit = self.n if first_dup == _U64_MAX else first_dup
        hist_d, present_d, scal_d = self.hist_distance(skip, skip_limit)
        if self.dist is not None:
            self.dist.allreduce_hist(hist_d, present_d, scal_d, fused=self._ws["k1buf"][:self.D + _capi.N_SCALARS])
        D = self.D
        hbuf = self._ws["k1buf"][:D + _capi.N_SCALARS + ((D + 31) // 32 + 1) // 2].cpu().numpy()
        hist = hbuf[:D]
        scal = hbuf[D:D + _capi.N_SCALARS]
        present = hbuf[D + _capi.N_SCALARS:].view(np.uint32)[:(D + 31) // 32]
        if int(scal[_capi.S_OFFGRID]) != 0:
            raise ValueError("%d in-range intra contacts have a distance that is not a multiple of the resolution %d "
                             "(or beyond the fragment list); only fixed-size bins on a common grid are supported"
                             % (int(scal[_capi.S_OFFGRID]), res))
        N = int(scal[_capi.S_INTRA_INRANGE_SUM])
        obsInterAllCount = int(scal[_capi.S_INTER_ALL_COUNT])
        obsInterAllSum = int(scal[_capi.S_INTER_ALL_SUM])
        obsIntraAllSum = int(scal[_capi.S_INTRA_ALL_SUM])
        max_count = int(scal[_capi.S_MAX_COUNT])
        t1 = time.perf_counter()
        # ---- host: bins, possible pairs, probabilities, spline fit ----
        if present.any():  # a distance whose counts sum to zero still counts as seen (:434-436): rare
            pres_bits = np.unpackbits(present.view(np.uint8), bitorder="little")[:D].astype(bool)
            seen = np.nonzero((hist != 0) | pres_bits)[0]
        else:
            seen = np.nonzero(hist)[0]
        dists = (seen * res).astype(np.int64)
        sums = hist[seen].astype(np.int64)
        bins = make_bins(lib, dists, sums, st.noOfBins, N)
        dec = None
        if passNo > 1 and outl is not None and bins["n"] > 0:
            dec = self.outlier_bin_decrements(outl, bins["ub"])
            if self.dist is not None:
                dec = self.dist.allreduce_small(dec)
        fp = frag_pairs(lib, self.frags, st, bins, dec)
        x, y = calculate_probabilities(bins, N)
        out = dict(passNo=passNo, N=N, dists=dists, sums=sums, bins=bins, x=x, y=y, x_bins=x, y_bins=y,
                   observedInterAllCount=obsInterAllCount, observedInterAllSum=obsInterAllSum,
                   observedIntraAllSum=obsIntraAllSum, observedIntraInRangeLines=int(scal[_capi.S_INTRA_INRANGE_LINES]),
                   observedIntraAllLines=int(scal[_capi.S_INTRA_ALL_LINES]), max_count=max_count, **fp)
        interChrProb = 1.0 / obsInterAllCount if obsInterAllCount > 0 else 0.0   # :669-672
        out["interChrProb"] = interChrProb
        lut = None
        if not st.interOnly:
            xs, ys, tck = fit_spline(x, y)
            splineX = dists[(dists >= min(xs)) & (dists <= max(xs))]
            out.update(x=xs, y=ys, tck=tck, splineX=splineX)
            t2 = time.perf_counter()
            table, lut = self.spline_table(tck, splineX, min(xs), max(xs))
            out["table_dev"] = table
        else:
            t2 = time.perf_counter()
        # ---- T (fithic/fithic.py:1128-1163) ----
        if st.allReg:
            T = out["possibleIntraInRangeCount"] + obsInterAllCount
        elif st.interOnly:
            T = obsInterAllCount
        else:
            T = out["possibleIntraInRangeCount"]
        out["T"] = T
        # ---- K3 ----
        thres = (1.0 / T) if T != 0 else float("inf")
        out["outlierThres"] = thres
        p, e = self.pvalues(lut, N, obsInterAllSum, interChrProb, max_count, outl, thres, outl_stats, pvalue_chunks,
                            after_chunk)
        if after_pvalues is not None:
            after_pvalues(p, e)  # e.g. start the device->host copy of p and ExpCC while K4 runs
        # ---- K4 ----
        if self.dist is not None:
            q = self.dist.global_bh(self, p, float(T))
        else:
            q = self.bh_qvalues(p, float(T))
        out.update(p=p, q=q, expcc=e)
        ev["k1_and_d2h"] = t1 - t0     # K1 launch + wait + histogram D2H (includes the device time of K1)
        ev["host_bins_fit"] = t2 - t1  # make_bins + frag_pairs + probabilities + scipy spline fit
        self.timings[passNo] = ev
        return out

    # ------------------------------------------------------------------------------------------------------------
    # K2
    MAX_CHR_RUNS = 256  # more chromosome runs than this: the dense chrs array is uploaded instead

    HOST_PAVA_MIN_POINTS = 4096  # above this the pooling runs on the host (see csrc/spline.cu)

    def spline_table(self, tck, splineX, xmin, xmax):
        t, c, k = tck
        assert k == 3
        D = self.distance_slots()
        m = int(len(splineX))
        if m == 0:
            raise ValueError("no observed distance falls inside the fitted range")
        host = np.concatenate([np.asarray(t, np.float64), np.asarray(c, np.float64)])
        tc = torch.from_numpy(host).to(self.device)
        sx = torch.from_numpy(np.ascontiguousarray(splineX, dtype=np.int64)).to(self.device)
        nt = len(t)
        table = self._tensor("table", m, torch.float64)
        lut = self._tensor("lut", D, torch.float64)
        res = self.grid
        if m < self.HOST_PAVA_MIN_POINTS:
            wsb = int(self.lib.fhc_spline_workspace_bytes(m))
            ws = self._buf("spline_ws", wsb)
            check(self.lib.fhc_spline_table(dptr(tc[:nt]), dptr(tc[nt:]), nt, dptr(sx), m, float(xmin), float(xmax), res,
                                            dptr(table), dptr(lut), D, dptr(ws), wsb, self._stream()))
        else:
            check(self.lib.fhc_spline_eval(dptr(tc[:nt]), dptr(tc[nt:]), nt, dptr(sx), m, dptr(table), self._stream()))
            y = table.cpu().numpy()
            check(self.lib.fhc_host_antitonic(dptr(y), m))
            table.copy_(torch.from_numpy(y))
            check(self.lib.fhc_spline_lut(dptr(sx), dptr(table), m, float(xmin), float(xmax), res, dptr(lut), D,
                                          self._stream()))
        self._keep = (tc, sx)  # keep the small inputs alive until the stream has consumed them
        return table, lut

    # lbeta table
    def lbeta_table(self, name, N, max_count):
        ntab = int(min(max(max_count, 1), min(N, (1 << 22) - 1)) + 1)
        tab = self._tensor(name, ntab, torch.float64)
        if os.environ.get("FHC_LBETA_TABLE") == "host":
            # the table from the C library's log (scipy's own) instead of the device kernel's correctly
            # rounded one: follows scipy in the rare entries where the two logs differ.  Opt-in until timed on the GPU.
            host = np.empty(ntab, dtype=np.float64)
            check(self.lib.fhc_host_lbeta_table(int(N), dptr(host), ntab, int(os.environ.get("FHC_HOST_THREADS", "8"))))
            tab.copy_(torch.from_numpy(host))
            return tab, ntab
        check(self.lib.fhc_lbeta_table(int(N), dptr(tab), ntab, self._stream()))
        return tab, ntab

    # K3  (fit_Spline per-line loop, fithic/fithic.py:1017-1123)
    def pvalues(self, lut, N_intra, N_inter, interChrProb, max_count, outl=None, outl_thres=0.0, outl_stats=None,
                nchunks=1, after_chunk=None):
        """nchunks > 1 launches K3 once per contiguous slice of the contacts and calls after_chunk(lo, hi, p, e) after
        each launch, so that a caller can start moving finished slices to the host while the next one is computed."""
        st = self.st
        mid1, mid2, cnt, chrs = self.contacts
        n = self.n
        p = self._tensor("p", n, torch.float64)
        e = self._tensor("expcc", n, torch.float64)
        tab_a = tab_b = None
        nta = ntb = 0
        if not st.interOnly:
            tab_a, nta = self.lbeta_table("lbeta_intra", N_intra, max_count)
        if st.interOnly or st.allReg:
            tab_b, ntb = self.lbeta_table("lbeta_inter", N_inter, max_count)
        bias = bmid = boff = None
        nchr = 0
        if self._bias_dev is not None:
            bias, bmid, boff = self._bias_dev
            nchr = boff.numel() - 1
        step = n if nchunks <= 1 else max(((n + nchunks - 1) // nchunks + 4095) // 4096 * 4096, 4096)
        wsb = int(self.lib.fhc_pvalues_workspace_bytes(min(step, n), max(nta, ntb)))
        ws = self._buf("pval_ws", wsb)  # work lists of the contacts that need an iterative evaluation
        lo = 0
        while True:
            hi = min(lo + step, n)
            o = None if outl is None else outl[lo:hi]
            check(self.lib.fhc_pvalues(st.mode, dptr(mid1[lo:hi]), dptr(mid2[lo:hi]), dptr(cnt[lo:hi]), dptr(chrs[lo:hi]),
                                       hi - lo, dptr(bias), dptr(bmid), dptr(boff), nchr,
                                       getattr(self, "_bias_sparse", 0), self.grid, st.L, st.U,
                                       dptr(lut), self.D if lut is not None else 0, int(N_intra), int(N_inter),
                                       float(interChrProb), float(st.biasLowerBound), float(st.biasUpperBound), dptr(tab_a),
                                       nta, dptr(tab_b), ntb, dptr(o), lo, float(outl_thres), dptr(outl_stats), dptr(p[lo:hi]),
                                       dptr(e[lo:hi]), dptr(ws), wsb, self._stream()))
            if after_chunk is not None:
                after_chunk(lo, hi, p, e)
            lo = hi
            if lo >= n:
                break
        return p, e

    # K4  (myStats.benjamini_hochberg_correction, fithic/myStats.py:24-48)
    def bh_qvalues(self, p, T, rank_offset=0, carry_in=0.0, carry_out=None, n_sorted_out=None, q=None):
        n = p.numel()
        if q is None:
            q = self._tensor("q", n, torch.float64)
        wsb = int(self.lib.fhc_bh_workspace_bytes(n))
        ws = self._buf("bh_ws", wsb)
        # one host sync after the compaction: only the radix passes the number of ranked keys needs are launched
        ranked = ctypes.c_int64(0)
        check(self.lib.fhc_bh_qvalues_hostcount(dptr(p), n, float(T), int(rank_offset), float(carry_in), dptr(q),
                                                dptr(carry_out), dptr(n_sorted_out), ctypes.byref(ranked), dptr(ws), wsb,
                                                self._stream()))
        self.last_ranked = int(ranked.value)
        return q

    # K5  (makeBinsFromInteractions outlier decrements, fithic/fithic.py:528-548)
    def outlier_bin_decrements(self, outl, bin_ub):
        mid1, mid2, _, _ = self.contacts
        nb = len(bin_ub)
        ub = torch.from_numpy(np.ascontiguousarray(bin_ub, dtype=np.int64)).to(self.device)
        dec = self._tensor("outl_dec", nb, torch.int64)
        check(self.lib.fhc_outlier_bin_decrements(dptr(mid1), dptr(mid2), dptr(outl), self.n, dptr(ub), nb, dptr(dec),
                                                  self._stream()))
        return dec.cpu().numpy()

    def new_outlier_state(self):
        outl = torch.zeros(max(self.n, 1), dtype=torch.uint8, device=self.device)[:self.n]
        stats = torch.tensor([0, -1], dtype=torch.int64, device=self.device)  # {flagged, first duplicate = UINT64_MAX}
        return outl, stats

    # ------------------------------------------------------------------------------------------------------------
    def run(self):
        """All spline passes (fithic/fithic.py:317-376).  Returns the list of per-pass dicts (device tensors p, q, expcc);
        the outlier multiplicities of the last pass stay in self.outl / self.outl_stats."""
        outl, stats = self.new_outlier_state()
        results = []
        for passNo in range(1, self.st.noOfPasses + 1):
            if passNo > 1 and self.st.interOnly:
                break  # :349-351
            results.append(self.run_pass(passNo, outl, stats))
        self.outl, self.outl_stats = outl, stats
        return results


# ----------------------------------------------------------------------------------------------------------------
# host stages between K1 and K2 (O(D) work)
# ----------------------------------------------------------------------------------------------------------------
def make_bins(lib, dists, sums, noOfBins, N):
    """makeBinsFromInteractions (fithic/fithic.py:463-553) through the C helper; returns dict of arrays."""
    dists = np.ascontiguousarray(dists, dtype=np.int64)
    sums = np.ascontiguousarray(sums, dtype=np.int64)
    lb = np.zeros(noOfBins, dtype=np.int64)
    ub = np.zeros(noOfBins, dtype=np.int64)
    sc = np.zeros(noOfBins, dtype=np.int64)
    nb = check(lib.fhc_host_make_bins(dptr(dists), dptr(sums), len(dists), int(noOfBins), int(N), dptr(lb), dptr(ub),
                                      dptr(sc)))
    return dict(n=nb, lb=lb[:nb].copy(), ub=ub[:nb].copy(), sumcc=sc[:nb].copy())


def frag_pairs(lib, frags, st, bins, dec=None):
    """generate_FragPairs (fithic/fithic.py:596-689 fixed-size bins, :691-778 restriction fragments).  Mutates `bins`
    (adds pairs = `[1]`, pairs7 = `[7]`, sumdist = `[3]`; the two pair counts differ only for restriction fragments)."""
    order = sorted(range(len(frags.chroms)), key=lambda i: frags.chroms[i])  # sorted chromosome NAMES (:606)
    order = [i for i in order if frags.n_mappable[i] > 0]
    nb = bins["n"]
    pairs = np.zeros(max(nb, 1), dtype=np.int64)
    if dec is not None:
        pairs[:nb] -= np.asarray(dec, dtype=np.int64)[:nb]
    sumdist = np.zeros(max(nb, 1), dtype=np.float64)
    if st.resolution == 0:
        mids = np.ascontiguousarray(np.concatenate([np.asarray(frags.mids[i], dtype=np.int64) for i in order])
                                    if order else np.zeros(0, dtype=np.int64))
        off = np.zeros(len(order) + 1, dtype=np.int64)
        np.cumsum([len(frags.mids[i]) for i in order], out=off[1:])
        pairs7 = pairs.copy()  # the outlier decrements hit [1] and [7] alike (:544-545)
        totals = np.zeros(5, dtype=np.int64)
        check(lib.fhc_host_frag_pairs_varsize(dptr(mids), dptr(off), len(order), st.L, st.U, dptr(bins["lb"]), dptr(bins["ub"]),
                                              nb, dptr(pairs), dptr(pairs7), dptr(sumdist), dptr(totals)))
        bins["pairs"], bins["pairs7"], bins["sumdist"] = pairs[:nb], pairs7[:nb], sumdist[:nb]
        return dict(possibleIntraInRangeCount=int(totals[0]), possibleIntraAllCount=int(totals[1]),
                    possibleInterAllCount=totals[2] / 2, noOfFrags=int(totals[3]), maxPossibleGenomicDist=int(totals[4]))
    chr_n = np.ascontiguousarray(frags.n_mappable[order], dtype=np.int64)
    chr_mm = np.ascontiguousarray(frags.max_mid[order], dtype=np.int64)
    totals = np.zeros(4, dtype=np.int64)
    check(lib.fhc_host_frag_pairs(dptr(chr_n), dptr(chr_mm), len(order), int(st.resolution), st.L, st.U,
                                  dptr(bins["lb"]), dptr(bins["ub"]), nb, dptr(pairs), dptr(sumdist), dptr(totals)))
    bins["pairs"] = pairs[:nb]
    bins["pairs7"] = bins["pairs"]
    bins["sumdist"] = sumdist[:nb]
    return dict(possibleIntraInRangeCount=int(totals[0]), possibleIntraAllCount=totals[1] / 2,
                possibleInterAllCount=totals[2] / 2, noOfFrags=int(totals[3]))


def calculate_probabilities(bins, N):
    """calculateProbabilities (fithic/fithic.py:869-908): x = avgDist, y = avgCC per bin (vectorised, same IEEE ops)."""
    nb = bins["n"]
    pairs = bins["pairs"].astype(np.float64)     # [1] (:877)
    pairs7 = bins["pairs7"].astype(np.float64)   # [7] (:885)
    sumcc = bins["sumcc"].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = np.where((bins["pairs"] > 0) & (N > 0), (1.0 * sumcc / pairs) / float(N) if N > 0 else 0.0, 0.0)
        x = np.where(bins["pairs7"] != 0, 1000000.0 * (bins["sumdist"] / pairs7), 0.0)
    return [float(v) for v in x[:nb]], [float(v) for v in y[:nb]]


def fit_spline(x, y):
    """fit_Spline fit stage (fithic/fithic.py:936-951): sort by x, require strictly increasing x, cubic
    UnivariateSpline with s = min(y)^2 (scipy FITPACK on <= noOfBins points; the per-distance evaluation is K2)."""
    from scipy.interpolate import UnivariateSpline
    y = [f for _, f in sorted(zip(x, y), key=lambda pair: pair[0])]
    x = sorted(x)
    for i in range(1, len(x)):
        if x[i] <= x[i - 1]:
            print("ERROR in spline fitting. Distances do not decrease across bins. Ensure interaction file is correct.")
            print("Avg. distance of bin(i-1)... %s" % x[i - 1])
            print("Avg. distance of bin(i)... %s" % x[i])
            raise SystemExit(2)
    ius = UnivariateSpline(x, y, s=min(y) * min(y))
    t, c, k = ius._eval_args
    return x, y, (np.asarray(t, np.float64), np.asarray(c, np.float64), int(k))
