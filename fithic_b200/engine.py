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
    biasUpperBound: float = 2  # (an int like the reference's default: the log prints it)
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


def host_threads():
    """Host threads of this rank for the stages between the kernels: FHC_HOST_THREADS, else the cores this rank can call
    its own, at most 8.  Under torchrun that is its share of the node's cores MINUS ONE: the pool's workers spin while a
    pass runs, and a node whose cores are all spinning has none left for the other threads of the ranks (measured on a
    32-core box with 8 GPUs: 8 threads per rank 6.4 ms per pass, 4 threads 2.02 ms, 3 threads 1.97 ms)."""
    env = os.environ.get("FHC_HOST_THREADS")
    if env:
        return max(1, min(64, int(env)))
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 4
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    if local > 1:
        if os.environ.get("FHC_CORES_PINNED") == "1":  # parallel.DistCtx gave this rank its own cores (FHC_PIN_CORES=1)
            return max(1, min(8, cores - 1))
        return max(1, min(8, cores // local - 1))
    return max(1, min(8, cores))


class PassResult(dict):
    """The dict run_pass returns.  `table_dev` (the spline table at the observed distances on the device) is uploaded on
    first use: K3 only needs the dense lookup table, and the native host stage leaves the table itself on the host."""

    def __missing__(self, key):
        if key == "table_dev" and "table" in self and getattr(self, "_device", None) is not None:
            t = torch.from_numpy(np.ascontiguousarray(self["table"])).to(self._device)
            self[key] = t
            return t
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (key == "table_dev" and dict.__contains__(self, "table")
                                                and getattr(self, "_device", None) is not None)


class _HostStage:
    """Buffers and the fhc_stage_io block of fhc_host_stage (csrc/hoststage.cu) for one engine: pinned staging buffers for
    K1's histogram, the lookup table and the lbeta tables; per-pass output arrays are fresh numpy arrays (they end up in
    the pass's result dict)."""

    LBETA_CAP0 = 1 << 14

    def __init__(self, eng):
        st, frags = eng.st, eng.frags
        order = sorted(range(len(frags.chroms)), key=lambda i: frags.chroms[i])  # sorted chromosome NAMES (:606)
        order = [i for i in order if frags.n_mappable[i] > 0]
        self.chr_n = np.ascontiguousarray(frags.n_mappable[order], dtype=np.int64)
        self.chr_mm = np.ascontiguousarray(frags.max_mid[order], dtype=np.int64)
        self.nthreads = host_threads()
        io = self.io = _capi.StageIO()
        io.grid = eng.grid
        io.noOfBins = int(st.noOfBins)
        io.L, io.U = st.L, st.U
        io.chr_n, io.chr_maxmid, io.nchr = self.chr_n.ctypes.data, self.chr_mm.ctypes.data, len(order)
        io.want_spline = 0 if st.interOnly else 1
        io.nthreads = self.nthreads
        self.D = 0
        self.event = None
        self.lbeta = [None, None]  # pinned tensors

    def ensure(self, D):
        if D == self.D:
            return
        self.D = D
        nwords = (D + 31) // 32
        self.k1 = torch.empty(D + _capi.N_SCALARS + 64 + (nwords + 1) // 2, dtype=torch.int64).pin_memory()
        self.k1_np = self.k1.numpy()
        self.lut = torch.empty(D, dtype=torch.float64).pin_memory()
        self.io.k1buf = self.k1.data_ptr()
        self.io.D = D
        self.io.lut = self.lut.data_ptr()

    def ensure_lbeta(self, which, cap):
        cur = self.lbeta[which]
        if cur is None or cur.numel() < cap:
            cur = self.lbeta[which] = torch.empty(int(cap), dtype=torch.float64).pin_memory()
        self.io.lbeta_tab[which] = cur.data_ptr()
        self.io.lbeta_cap[which] = cur.numel()
        return cur

    def new_outputs(self):
        """The stage writes into two blocks that live as long as the engine (fresh arrays would cost a page fault per
        4 kB on the critical path: 0.3 ms per pass at 5 kb); run_pass copies what goes into the result dict out of them
        after it has launched K3 and K4, while the GPU is busy."""
        D, nb = self.D, int(self.io.noOfBins)
        io = self.io
        if getattr(self, "_blk_shape", None) == (D, nb):
            return
        self._blk_shape = (D, nb)
        self.i64 = i64 = np.zeros(3 * D + 4 * nb, dtype=np.int64)
        self.f64 = f64 = np.zeros(D + 5 * nb + 2 * (nb + 4), dtype=np.float64)
        b = i64.ctypes.data
        io.dists, io.sums, io.splineX = b, b + 8 * D, b + 16 * D
        b += 24 * D
        io.bin_lb, io.bin_ub, io.bin_sumcc, io.bin_pairs = b, b + 8 * nb, b + 16 * nb, b + 24 * nb
        b = f64.ctypes.data
        io.table = b
        b += 8 * D
        io.bin_sumdist, io.x_bins, io.y_bins, io.xs, io.ys = (b + 8 * nb * k for k in range(5))
        b += 40 * nb
        io.t, io.c = b, b + 8 * (nb + 4)

    def views(self):
        D, nb = self.D, int(self.io.noOfBins)
        i64, f64, io = self.i64, self.f64, self.io
        n, ns, m, nt = int(io.nb), int(io.nseen), int(io.m), int(io.nt)
        o = 3 * D
        v = dict(dists=i64[:ns], sums=i64[D:D + ns], splineX=i64[2 * D:2 * D + m],
                 lb=i64[o:o + n], ub=i64[o + nb:o + nb + n], sumcc=i64[o + 2 * nb:o + 2 * nb + n],
                 pairs=i64[o + 3 * nb:o + 3 * nb + n], table=f64[:m])
        o = D
        for k, name in enumerate(("sumdist", "x_bins", "y_bins", "xs", "ys")):
            v[name] = f64[o + k * nb:o + k * nb + n]
        o = D + 5 * nb
        v["t"], v["c"] = f64[o:o + nt], f64[o + nb + 4:o + nb + 4 + nt]
        return v


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
        self.chr_runs_dev = None
        self.n = 0

    # ------------------------------------------------------------------------------------------------------------
    def _stream(self):
        ps = getattr(self, "_pass_stream", None)  # inside run_pass: looked up once per pass
        if ps is not None:
            return ps
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
    def set_contacts_device(self, mid1, mid2, cnt, chrs, chr_runs=None):
        """Adopt device-resident int32 tensors (chrs holds the uint32 bit pattern).  chr_runs = (values uint32 [r], lengths
        int64 [r]), the run-length form of chrs (see Contacts.chr_runs): with it K1 and K3 never read a chrs array, and
        `chrs` may be None."""
        n = mid1.numel()
        for t in (mid1, mid2, cnt) + ((chrs,) if chrs is not None else ()):
            assert t.is_cuda and t.dtype == torch.int32 and t.numel() == n and t.is_contiguous()
        self.n = n
        self.D = None
        self._own_contacts = False
        self.chr_runs_dev = None
        self._chr_runs_host = None
        if chr_runs is not None and 1 <= len(chr_runs[0]) <= _capi.MAX_CHR_RUNS and n > 0 \
                and os.environ.get("FHC_CHR_RUNS", "1") != "0":
            vals = np.ascontiguousarray(chr_runs[0], dtype=np.uint32)
            starts = np.zeros(len(vals) + 1, dtype=np.int64)
            np.cumsum(np.asarray(chr_runs[1], dtype=np.int64), out=starts[1:])
            assert int(starts[-1]) == n, "chr_runs do not cover the contacts"
            self._chr_runs_host = (vals, starts)
            self.chr_runs_dev = (torch.from_numpy(starts).to(self.device), torch.from_numpy(vals.view(np.int32)).to(self.device),
                                 len(vals))
        elif chrs is None:
            if chr_runs is None:
                raise ValueError("need the chrs array or its run-length form")
            chrs = self._expand_runs(chr_runs, n)
        self.contacts = (mid1, mid2, cnt, chrs)

    def _expand_runs(self, chr_runs, n, into=None):
        """chrs array from its run-length form (one fill per run; torch.repeat_interleave parallelises over the runs, not over
        the output: 30 ms for 24 runs of 12 M)."""
        d = into if into is not None else torch.empty(n, dtype=torch.int32, device=self.device)
        pos = 0
        for v, ln in zip(np.asarray(chr_runs[0]).view(np.int32).tolist(), np.asarray(chr_runs[1]).tolist()):
            d[pos:pos + ln].fill_(v)
            pos += ln
        assert pos == n, "chr_runs do not cover the contacts"
        return d

    def chrs_array(self):
        """The per-line chrs array on the device; built from the runs on first use when the contacts came without one (only
        the tile-phased K3 and callers outside the hot path ask for it)."""
        mid1, mid2, cnt, chrs = self.contacts
        if chrs is None:
            vals, starts = self._chr_runs_host
            chrs = self._expand_runs((vals, np.diff(starts)), self.n)
            self.contacts = (mid1, mid2, cnt, chrs)
        return chrs

    def upload_contacts(self, c, non_blocking=False):
        """Host SoA -> HBM (the only per-contact host->device traffic of a run: 12 B per contact when the chromosome ids
        come as runs, 16 B otherwise)."""
        ts = []
        reuse = self.contacts if (self.contacts is not None and getattr(self, "_own_contacts", False)
                                  and self.n == len(c)) else None
        runs = c.chr_runs if (c.chr_runs is not None and 1 <= len(c.chr_runs[0]) <= _capi.MAX_CHR_RUNS
                              and os.environ.get("FHC_CHR_RUNS", "1") != "0") else None
        for j, a in enumerate((c.mid1, c.mid2, c.cnt, c.chrs.view(np.int32) if runs is None else None)):
            if a is None:
                ts.append(None)  # chromosome ids from their run-length form: nothing over the link, nothing in HBM
                continue
            h = torch.from_numpy(np.ascontiguousarray(a))
            if reuse is not None and reuse[j] is not None:  # same size as the previous upload: no new device allocation
                reuse[j].copy_(h, non_blocking=non_blocking)
                ts.append(reuse[j])
            else:
                ts.append(h.to(self.device, non_blocking=non_blocking))
        self.set_contacts_device(*ts, chr_runs=runs)
        self._own_contacts = True

    def distance_slots(self):
        """D = number of distance slots: every |mid1 - mid2| of an intra line is < D * res."""
        if self.D is None:
            mid1, mid2, _, _ = self.contacts
            if self.n == 0:
                mx = 0
            else:
                rng = self._tensor("mid_range", 2, torch.int64)
                check(self.lib.fhc_mid_range(dptr(mid1), dptr(mid2), self.n, dptr(rng), self._stream()))
                lo, hi = rng.cpu().tolist()
                mx = int(hi) - int(lo)
            if len(self.frags.max_mid):
                mx = max(mx, int(self.frags.max_mid.max()))
            if self.st.resolution == 0 and self.st.U >= 0:
                mx = min(mx, self.st.U)  # on the 1 bp grid only in-range distances need a slot
            self.D = mx // self.grid + 2
            if self.st.resolution == 0 and self.D > (1 << 27):
                # the reference keeps its distances in a dict; here every bp of the longest distance is a slot of the
                # histogram and of the lookup table (8 B each, plus their host copies per pass)
                import warnings
                warnings.warn("restriction-fragment mode without -U: the distance axis has %d one-bp slots (%.1f GB per "
                              "table, moved to the host once per pass); give -U to bound it" % (self.D, self.D * 8 / 1e9))
            if self.dist is not None:
                self.D = self.dist.max_int(self.D)  # the histogram is all-reduced: every rank needs the same length
        return self.D

    # ------------------------------------------------------------------------------------------------------------
    # K1  (read_Interactions, fithic/fithic.py:389-454)
    def _rank_slots(self):
        """Slots behind K1's totals, one per rank for the largest count (plus one of padding when that makes the summed
        part [hist | totals | slots] an even number of words: the library's own all-reduce moves 16-byte words)."""
        if self.dist is None:
            return 0
        w = self.dist.world
        return w + ((self.distance_slots() + _capi.N_SCALARS + w) & 1)

    def hist_distance(self, skip=None, skip_limit=-1):
        D = self.distance_slots()
        mid1, mid2, cnt, chrs = self.contacts
        # one buffer [hist | totals | one slot per rank for the largest count | seen bitmap]: a single all-reduce(sum) between
        # GPUs and a single device->host copy per pass
        nwords = (D + 31) // 32
        slots = self._rank_slots()
        my = self.dist.rank if self.dist is not None else 0
        ns = _capi.N_SCALARS + slots
        buf = self._tensor("k1buf", D + ns + (nwords + 1) // 2, torch.int64)
        hist = buf[:D]
        scal = buf[D:D + ns]
        present = buf[D + ns:].view(torch.int32)[:nwords]
        rs, rv, nruns = self.chr_runs_dev if self.chr_runs_dev is not None else (None, None, 0)
        check(self.lib.fhc_hist_distance(dptr(mid1), dptr(mid2), dptr(cnt), None if nruns else dptr(self.chrs_array()), dptr(rs), dptr(rv),
                                         nruns, dptr(skip), int(skip_limit), self.n, self.st.L, self.st.U, self.grid,
                                         dptr(hist), dptr(present), D, dptr(scal), slots, my, self._stream()))
        return hist, present, scal

    # ------------------------------------------------------------------------------------------------------------
    # the native host stage covers fixed-size bins with a distance axis of at most this many slots (-r 0 on its 1 bp grid
    # and anything larger keep the staged path: evaluation and lookup table on the device)
    NATIVE_STAGE_MAX_SLOTS = 1 << 20
    PREPASS_AUTO_MAX = 120_000_000  # FHC_PREPASS=auto: contacts per GPU up to which the pre-pass hides behind the host stage
    # ... and what a larger shard gets: the lines that fit into the gap (B200: host stage ~0.55 ms, q fill 5.7 TB/s, pre-pass
    # 170 M lines per ms); fewer than PREPASS_PARTIAL_MIN lines are not worth K3's second set of launches (measured: 22 M of
    # 300 M lines on one GPU 9.29 against 9.21 ms per pass; 58 M of 150 M lines on each of two GPUs 4.84 against 5.00 ms)
    PREPASS_GAP_MS = 0.55
    PREPASS_FILL_BYTES_PER_MS = 5.7e9
    PREPASS_LINES_PER_MS = 170e6
    PREPASS_PARTIAL_MIN = 32_000_000

    def run_pass(self, passNo, outl=None, outl_stats=None, after_pvalues=None, pvalue_chunks=1, after_chunk=None):
        """One spline pass.  Returns a dict with host-side tables and device tensors p, q, expcc."""
        self._pass_stream = None
        self._pass_stream = self._stream()
        try:
            return self._run_pass(passNo, outl, outl_stats, after_pvalues, pvalue_chunks, after_chunk)
        finally:
            self._pass_stream = None

    def _run_pass(self, passNo, outl, outl_stats, after_pvalues, pvalue_chunks, after_chunk):
        st = self.st
        t0 = time.perf_counter()
        # ---- K1 ----
        skip, skip_limit = None, -1
        if passNo > 1 and outl is not None:
            skip = outl
            first_dup = int(outl_stats[1].item()) & _U64_MAX
            if self.dist is not None and passNo > 2:
                # The reference stops skipping outlier lines after the first duplicated entry of its sorted outlier list
                # (from pass 3 on, fithic/fithic.py:408-412), a position in FILE order: the ranks agree on the smallest
                # global line number of a first duplicate and translate it back into their own lines.
                skip_limit = self._global_skip_limit(first_dup)
            else:
                skip_limit = self.n if first_dup == _U64_MAX else first_dup
        hist_d, present_d, scal_d = self.hist_distance(skip, skip_limit)
        native = (st.resolution > 0 and self.D <= self.NATIVE_STAGE_MAX_SLOTS
                  and os.environ.get("FHC_HOST_STAGE", "native") != "legacy")
        if self.dist is not None:  # exchange 1: [hist | totals | rank slots] summed over the GPUs in one collective
            self.dist.allreduce_k1(self._ws["k1buf"][:self.D + scal_d.numel()])
        tables = self._tables_native if native else self._tables_legacy
        out, lut, lbeta, ev = tables(passNo, outl if passNo > 1 else None, t0)
        N, obsInterAllCount, obsInterAllSum = out["N"], out["observedInterAllCount"], out["observedInterAllSum"]
        interChrProb = 1.0 / obsInterAllCount if obsInterAllCount > 0 else 0.0   # :669-672
        out["interChrProb"] = interChrProb
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
        p, e = self.pvalues(lut, N, obsInterAllSum, interChrProb, out["max_count"], outl, thres, outl_stats, pvalue_chunks,
                            after_chunk, lbeta=lbeta, use_prepass=native)
        if after_pvalues is not None:
            after_pvalues(p, e)  # e.g. start the device->host copy of p and ExpCC while K4 runs
        fin = getattr(out, "_finish", None)
        if fin is not None:  # the native host stage's report on bins and fit leaves its buffers now, while K3 runs
            fin()
            out._finish = None
        # ---- K4 ----
        pre_q = getattr(self, "_q_prefilled", False)  # _tables_native filled q with 1.0 while the host was fitting
        self._q_prefilled = False
        if self.dist is not None:
            q = self.dist.global_bh(self, p, float(T), q_prefilled=pre_q)
        else:
            q = self.bh_qvalues(p, float(T), q_prefilled=pre_q)
        out.update(p=p, q=q, expcc=e)
        self.timings[passNo] = ev
        return out

    def set_line_runs(self, starts, lengths):
        """Where this rank's lines sit in the whole file: local lines come as runs, run j = `lengths[j]` consecutive lines
        of the file starting at global line `starts[j]` (ascending).  Needed by multi-GPU runs with more than two spline
        passes (the reference's outlier skipping stalls at a position in file order) and by writers that restore file order."""
        starts, lengths = np.asarray(starts, dtype=np.int64), np.asarray(lengths, dtype=np.int64)
        assert len(starts) == len(lengths) and int(lengths.sum()) == self.n, "line runs do not cover this rank's contacts"
        assert np.all(starts[1:] >= starts[:-1] + lengths[:-1]), "line runs must be ascending and disjoint"
        self.line_runs = (starts, lengths)

    def digest(self, p, q):
        """Order-independent digest (two uint64 sums of hashes) of (file line, p bits, q bits) over this rank's lines
        (fhc_digest_lines); shards' digests add up mod 2^64.  Without set_line_runs() the local lines count as file lines."""
        runs = getattr(self, "line_runs", None)
        if runs is None:
            runs = (np.zeros(1, dtype=np.int64), np.array([self.n], dtype=np.int64))
        starts, lens = runs
        keep = lens > 0
        starts, lens = starts[keep], lens[keep]
        if len(starts) == 0:
            return 0, 0
        loc = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=loc[1:])
        rl = torch.from_numpy(loc).to(self.device)
        rg = torch.from_numpy(np.ascontiguousarray(starts)).to(self.device)
        out = torch.zeros(2, dtype=torch.int64, device=self.device)
        check(self.lib.fhc_digest_lines(dptr(p), dptr(q), self.n, dptr(rl), dptr(rg), len(lens), dptr(out), self._stream()))
        a, b = out.cpu().numpy().view(np.uint64).tolist()
        return int(a), int(b)

    def _global_skip_limit(self, first_dup_local):
        runs = getattr(self, "line_runs", None)
        if runs is None:
            raise ValueError("a multi-GPU run with more than 2 spline passes needs Engine.set_line_runs(): from pass 3 on the "
                             "reference stops skipping outliers at a position in file order")
        starts, lens = runs
        ends = np.cumsum(lens)
        big = (1 << 63) - 1
        g = big
        if first_dup_local != _U64_MAX:
            j = int(np.searchsorted(ends, first_dup_local, side="right"))
            g = int(starts[j] + (first_dup_local - (ends[j] - lens[j])))
        g = self.dist.min_int(g)
        if g == big:
            return self.n
        return int(np.clip(g - starts + 1, 0, lens).sum()) - 1  # local lines whose global number is <= g form a prefix

    def _pass_scalars(self, passNo, scal):
        if int(scal[_capi.S_OFFGRID]) != 0:
            raise ValueError("%d in-range intra contacts have a distance that is not a multiple of the resolution %d "
                             "(or beyond the fragment list); only fixed-size bins on a common grid are supported"
                             % (int(scal[_capi.S_OFFGRID]), self.grid))
        return PassResult(passNo=passNo, N=int(scal[_capi.S_INTRA_INRANGE_SUM]),
                          observedInterAllCount=int(scal[_capi.S_INTER_ALL_COUNT]),
                          observedInterAllSum=int(scal[_capi.S_INTER_ALL_SUM]),
                          observedIntraAllSum=int(scal[_capi.S_INTRA_ALL_SUM]),
                          observedIntraInRangeLines=int(scal[_capi.S_INTRA_INRANGE_LINES]),
                          observedIntraAllLines=int(scal[_capi.S_INTRA_ALL_LINES]), max_count=int(scal[_capi.S_MAX_COUNT]))

    def _tables_native(self, passNo, outl, t0):
        """Histogram -> lookup table through fhc_host_stage: one device->host copy of K1's buffer, one C call for bins,
        possible pairs, probabilities, spline fit, table and lbeta tables, and host->device copies of the lookup table and
        the lbeta tables on the pass's stream."""
        st, lib, D = self.st, self.lib, self.D
        hs = getattr(self, "_stage", None)
        if hs is None:
            hs = self._stage = _HostStage(self)
        hs.ensure(D)
        io = hs.io
        stream = self._stream()
        k1buf = self._ws["k1buf"]
        slots = self._rank_slots()
        io.n_rank_slots = slots
        shm = getattr(self.dist, "shm", None) if self.dist is not None else None
        if shm is not None:  # the ranks of the node share the possible-pair sums (fhc_host_stage phase 2)
            io.pairs_rank, io.pairs_world, io.shm = self.dist.rank, self.dist.world, shm
        else:
            io.pairs_rank, io.pairs_world, io.shm = 0, 1, None
        nk = D + _capi.N_SCALARS + slots
        check(lib.fhc_copy_async(hs.k1.data_ptr(), k1buf.data_ptr(), 8 * nk, stream))
        if hs.event is None:
            ev = ctypes.c_void_p()
            check(lib.fhc_event_create(ctypes.byref(ev)))
            hs.event = ev
        check(lib.fhc_event_record(hs.event, stream))
        # behind the copy, while the host bins and fits: q = 1.0 everywhere (K4 then only writes the lines that differ), and
        # the part of K3 that does not need the spline table
        self._q_prefilled = False
        if os.environ.get("FHC_Q_PREFILL", "1") != "0" and self.n > 0:
            check(lib.fhc_fill_f64(dptr(self._tensor("q", self.n, torch.float64)), self.n, 1.0, stream))
            self._q_prefilled = True
        self.prepass(passNo)
        lib.fhc_host_pool_prewarm(hs.nthreads)
        hs.new_outputs()
        use_host_lbeta = os.environ.get("FHC_LBETA_TABLE", "host") != "device"
        want = (not st.interOnly, st.interOnly or st.allReg)
        for w in (0, 1):
            if use_host_lbeta and want[w]:
                hs.ensure_lbeta(w, hs.LBETA_CAP0)
            else:
                io.lbeta_tab[w] = None
                io.lbeta_cap[w] = 0
        check(lib.fhc_event_synchronize(hs.event))  # the histogram is on the host (the pre-pass may still be running)
        scal = hs.k1_np[D:nk]
        if slots:
            scal[_capi.S_MAX_COUNT] = scal[_capi.N_SCALARS:].max()
        if int(scal[_capi.S_NONPOS_LINES]) != 0:  # distances seen only through lines with a count <= 0 (:434-436): rare
            nw = (D + 31) // 32
            if self.dist is not None:  # (every rank sees the same summed total, so every rank comes here)
                self.dist.or_present(k1buf[nk:].view(torch.int32)[:nw])
            check(lib.fhc_copy_async(hs.k1.data_ptr() + 8 * nk, k1buf.data_ptr() + 8 * nk, 4 * nw, stream))
            check(lib.fhc_stream_synchronize(stream))
        out = self._pass_scalars(passNo, scal)
        t1 = time.perf_counter()
        io.dec = None
        if outl is not None:
            check(lib.fhc_host_stage(ctypes.byref(io), 1))
            dec = None
            if io.nb > 0:
                dec = self.outlier_bin_decrements(outl, hs.views()["ub"])
                if self.dist is not None:
                    dec = self.dist.allreduce_small(dec)
                dec = np.ascontiguousarray(dec, dtype=np.int64)
                io.dec = dec.ctypes.data
            check(lib.fhc_host_stage(ctypes.byref(io), 6))
            io.dec = None
        else:
            check(lib.fhc_host_stage(ctypes.byref(io), 7))
        if io.status == 3:  # counts larger than the lbeta staging buffers: enlarge and fill them
            for w in (0, 1):
                if io.lbeta_cap[w] and io.lbeta_ntab[w] > io.lbeta_cap[w]:
                    hs.ensure_lbeta(w, int(io.lbeta_ntab[w]))
            check(lib.fhc_host_stage(ctypes.byref(io), 8))
        v = hs.views()
        if io.status == 1:
            i = int(io.bad_index)
            print("ERROR in spline fitting. Distances do not decrease across bins. Ensure interaction file is correct.")
            print("Avg. distance of bin(i-1)... %s" % v["xs"][i - 1])
            print("Avg. distance of bin(i)... %s" % v["xs"][i])
            raise SystemExit(2)
        if io.status == 2:
            raise ValueError("no observed distance falls inside the fitted range")
        if io.status == 4:
            raise ValueError("the spline fit needs more than 3 bins (got %d)" % int(io.nb))
        tot = io.totals
        out.update(possibleIntraInRangeCount=int(tot[0]), possibleIntraAllCount=tot[1] / 2, possibleInterAllCount=tot[2] / 2,
                   noOfFrags=int(tot[3]))
        # the tables K3 waits for leave first ...
        lut = None
        if not st.interOnly:
            lut = self._tensor("lut", D, torch.float64)
            check(lib.fhc_copy_async(lut.data_ptr(), hs.lut.data_ptr(), 8 * D, stream))
        lbeta = None
        if use_host_lbeta:
            lbeta = [None, 0, None, 0]
            for w, name in ((0, "lbeta_intra"), (1, "lbeta_inter")):
                if want[w]:
                    ntab = int(io.lbeta_ntab[w])
                    tab = self._tensor(name, ntab, torch.float64)
                    check(lib.fhc_copy_async(tab.data_ptr(), hs.lbeta[w].data_ptr(), 8 * ntab, stream))
                    lbeta[2 * w], lbeta[2 * w + 1] = tab, ntab
        nb, ier, calls = int(io.nb), int(io.ier), int(io.calls)

        # ... and what the pass reports about bins and fit is copied out of the stage's buffers once K3 is launched (run_pass)
        def finish():
            pairs = v["pairs"].copy()
            bins = dict(n=nb, lb=v["lb"].copy(), ub=v["ub"].copy(), sumcc=v["sumcc"].copy(), pairs=pairs, pairs7=pairs,
                        sumdist=v["sumdist"].copy())
            x_bins, y_bins = v["x_bins"].tolist(), v["y_bins"].tolist()
            out.update(bins=bins, x=x_bins, y=y_bins, x_bins=x_bins, y_bins=y_bins, dists=v["dists"].copy(),
                       sums=v["sums"].copy())
            if not st.interOnly:
                out.update(x=v["xs"].tolist(), y=v["ys"].tolist(), tck=(v["t"].copy(), v["c"].copy(), 3), spline_ier=ier,
                           spline_calls=calls, splineX=v["splineX"].copy(), table=v["table"].copy())
                out._device = self.device
        out._finish = finish
        t2 = time.perf_counter()
        tm = io.timings
        ev = {"k1_and_d2h": t1 - t0, "host_bins_fit": t2 - t1, "stage_bins": tm[0] * 1e-3, "stage_pairs_lbeta": tm[1] * 1e-3,
              "stage_fit": tm[2] * 1e-3, "stage_eval": tm[3] * 1e-3, "stage_antitonic": tm[4] * 1e-3, "stage_lut": tm[5] * 1e-3}
        return out, lut, lbeta, ev

    def _tables_legacy(self, passNo, outl, t0):
        """The same through the stage-by-stage entry points (restriction-fragment mode and very long distance axes):
        bins and possible pairs in C, the fit in C, evaluation and lookup table on the device."""
        st, lib, res, D = self.st, self.lib, self.grid, self.D
        slots = self._rank_slots()
        ns = _capi.N_SCALARS + slots
        nw = (D + 31) // 32
        if self.dist is not None:  # rare and cheap enough here: always OR the "distance seen" bitmaps
            self.dist.or_present(self._ws["k1buf"][D + ns:].view(torch.int32)[:nw])
        hbuf = self._ws["k1buf"][:D + ns + (nw + 1) // 2].cpu().numpy()
        hist = hbuf[:D]
        scal = hbuf[D:D + ns]
        if slots:
            scal[_capi.S_MAX_COUNT] = scal[_capi.N_SCALARS:].max()
        present = hbuf[D + ns:].view(np.uint32)[:nw]
        out = self._pass_scalars(passNo, scal)
        N = out["N"]
        t1 = time.perf_counter()
        # ---- host: bins, possible pairs, probabilities, spline fit ----
        if int(scal[_capi.S_NONPOS_LINES]) != 0:  # a distance whose counts sum to zero still counts as seen (:434-436): rare
            pres_bits = np.unpackbits(present.view(np.uint8), bitorder="little")[:D].astype(bool)
            seen = np.nonzero((hist != 0) | pres_bits)[0]
        else:
            seen = np.nonzero(hist)[0]
        dists = (seen * res).astype(np.int64)
        sums = hist[seen].astype(np.int64)
        bins = make_bins(lib, dists, sums, st.noOfBins, N)
        dec = None
        if outl is not None and bins["n"] > 0:
            dec = self.outlier_bin_decrements(outl, bins["ub"])
            if self.dist is not None:
                dec = self.dist.allreduce_small(dec)
        fp = frag_pairs(lib, self.frags, st, bins, dec, stream=self._stream())
        x, y = calculate_probabilities(bins, N)
        out.update(dists=dists, sums=sums, bins=bins, x=x, y=y, x_bins=x, y_bins=y, **fp)
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
        return out, lut, None, {"k1_and_d2h": t1 - t0, "host_bins_fit": t2 - t1}

    # ------------------------------------------------------------------------------------------------------------
    # K2
    MAX_CHR_RUNS = _capi.MAX_CHR_RUNS  # more chromosome runs than this: the dense chrs array is uploaded instead

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
        if os.environ.get("FHC_LBETA_TABLE", "host") != "device":
            # Default: the table from the C library's log -- the function scipy's cephes calls -- so that K3 follows scipy
            # also in the rare entries where that log is not correctly rounded (one ulp of lgam(N) = 4e-6 ... 8e-6 in p,
            # DESIGN.md section 2).  FHC_LBETA_TABLE=device runs the device kernel (correctly rounded log) instead.
            host = np.empty(ntab, dtype=np.float64)
            check(self.lib.fhc_host_lbeta_table(int(N), dptr(host), ntab, host_threads()))
            tab.copy_(torch.from_numpy(host))
            return tab, ntab
        check(self.lib.fhc_lbeta_table(int(N), dptr(tab), ntab, self._stream()))
        return tab, ntab

    def prepass(self, passNo):
        """fhc_pvalues_prepass over all contacts (work-list K3 only): bias products, line classes and distance slots, 12 B
        per contact.  Computed in the first pass of a run -- launched right behind K1, so it runs while the host bins and
        fits -- and reused by the later passes (it depends on nothing a pass changes)."""
        self._pre_live = False
        mode = os.environ.get("FHC_PREPASS", "auto")
        if mode == "0" or os.environ.get("FHC_PVAL_IMPL", "lists")[:1] == "t" or self.distance_slots() >= (1 << 30) \
                or self.n == 0:
            return
        # The pre-pass costs more instructions than it takes out of the front kernel (it re-reads the mid points and stores
        # 12 B per contact): it pays where it hides behind the host stage (about 0.55 ms, of which the q fill takes its
        # share) and in runs with several spline passes, which reuse it.  Shards up to ~120 M contacts (every multi-GPU run of
        # a whole genome) are pre-passed whole; of a larger shard only the first lines are, as many as fit into the gap --
        # K3 then runs once over those behind the pre-pass and once over the rest without.
        n_pre = self.n
        if mode == "auto" and self.n > self.PREPASS_AUTO_MAX and self.st.noOfPasses < 2:
            gap_ms = self.PREPASS_GAP_MS - self.n * 8 / self.PREPASS_FILL_BYTES_PER_MS
            n_pre = int(max(gap_ms, 0.0) * self.PREPASS_LINES_PER_MS) // 4096 * 4096
            if n_pre < self.PREPASS_PARTIAL_MIN:
                return
            n_pre = min(n_pre, self.n)
        forced = os.environ.get("FHC_PREPASS_LINES")  # tests: pre-pass exactly this many lines (rounded down to 4096)
        if forced:
            n_pre = min(int(forced) // 4096 * 4096, self.n)
            if n_pre <= 0:
                return
        self._pre_n = n_pre
        code = self._tensor("pre_code", n_pre, torch.int32)
        b12 = self._tensor("pre_b12", n_pre, torch.float64)
        key = (self.contacts[0].data_ptr(), self.n, n_pre, id(self._bias_dev))
        if passNo == 1 or getattr(self, "_pre_key", None) != key:
            st = self.st
            mid1, mid2, cnt, chrs = self.contacts
            rs, rv, nruns = self.chr_runs_dev if self.chr_runs_dev is not None else (None, None, 0)
            bias = bmid = boff = None
            nchr = 0
            if self._bias_dev is not None:
                bias, bmid, boff = self._bias_dev
                nchr = boff.numel() - 1
            check(self.lib.fhc_pvalues_prepass(st.mode, dptr(mid1), dptr(mid2), None if nruns else dptr(self.chrs_array()),
                                               dptr(rs), dptr(rv), nruns, n_pre, dptr(bias), dptr(bmid), dptr(boff), nchr,
                                               getattr(self, "_bias_sparse", 0), self.grid, st.L, st.U,
                                               float(st.biasLowerBound), float(st.biasUpperBound), 0, dptr(code), dptr(b12),
                                               self._stream()))
            self._pre_key = key
        self._pre_live = True

    # K3  (fit_Spline per-line loop, fithic/fithic.py:1017-1123)
    def pvalues(self, lut, N_intra, N_inter, interChrProb, max_count, outl=None, outl_thres=0.0, outl_stats=None,
                nchunks=1, after_chunk=None, lbeta=None, use_prepass=False):
        """nchunks > 1 launches K3 once per contiguous slice of the contacts and calls after_chunk(lo, hi, p, e) after
        each launch, so that a caller can start moving finished slices to the host while the next one is computed."""
        st = self.st
        mid1, mid2, cnt, chrs = self.contacts
        rs, rv, nruns = self.chr_runs_dev if self.chr_runs_dev is not None else (None, None, 0)
        if nruns and os.environ.get("FHC_PVAL_IMPL", "lists")[:1] != "t":
            chrs = None  # the work-list pipeline reads the runs; a tile inside one intra run needs no chromosome ids at all
        else:
            chrs = self.chrs_array()  # the tile-phased kernel reads the per-line array
            rs, rv, nruns = None, None, 0
        n = self.n
        pre_code = pre_b12 = None
        n_pre = 0
        if use_prepass and getattr(self, "_pre_live", False):
            n_pre = min(getattr(self, "_pre_n", n), n)
            pre_code, pre_b12 = self._ws["pre_code"][:n_pre], self._ws["pre_b12"][:n_pre]
        p = self._tensor("p", n, torch.float64)
        e = self._tensor("expcc", n, torch.float64)
        tab_a = tab_b = None
        nta = ntb = 0
        if lbeta is not None:  # tables of this pass already on their way to the device (fhc_host_stage)
            tab_a, nta, tab_b, ntb = lbeta
        else:
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
        # slices: the caller's chunks, cut once more where the pre-passed lines end (a slice is behind the pre-pass or not)
        cuts = sorted(set((list(range(0, n, step)) if n > 0 else [0]) + [n] + ([n_pre] if 0 < n_pre < n else [])))
        if len(cuts) < 2:
            cuts = [0, n]
        wsb = int(self.lib.fhc_pvalues_workspace_bytes(max(b - a for a, b in zip(cuts[:-1], cuts[1:])), max(nta, ntb)))
        ws = self._buf("pval_ws", wsb)  # work lists of the contacts that need an iterative evaluation
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            whole = lo == 0 and hi == n  # (a slice of a tensor costs microseconds: the usual single launch takes none)
            sl = (lambda t: t) if whole else (lambda t: t[lo:hi])
            o = None if outl is None else sl(outl)
            pre = pre_code is not None and hi <= n_pre
            check(self.lib.fhc_pvalues(st.mode, dptr(sl(mid1)), dptr(sl(mid2)), dptr(sl(cnt)),
                                       None if chrs is None else dptr(sl(chrs)), dptr(rs), dptr(rv), nruns,
                                       hi - lo, dptr(bias), dptr(bmid), dptr(boff), nchr,
                                       getattr(self, "_bias_sparse", 0), self.grid, st.L, st.U,
                                       dptr(lut), self.D if lut is not None else 0, int(N_intra), int(N_inter),
                                       float(interChrProb), float(st.biasLowerBound), float(st.biasUpperBound), dptr(tab_a),
                                       nta, dptr(tab_b), ntb, dptr(o), lo, float(outl_thres), dptr(outl_stats), dptr(sl(p)),
                                       dptr(sl(e)), dptr(pre_code[lo:hi]) if pre else None,
                                       dptr(pre_b12[lo:hi]) if pre else None, dptr(ws), wsb, self._stream()))
            if after_chunk is not None:
                after_chunk(lo, hi, p, e)
        return p, e

    # K4  (myStats.benjamini_hochberg_correction, fithic/myStats.py:24-48)
    def bh_qvalues(self, p, T, rank_offset=0, carry_in=0.0, carry_out=None, n_sorted_out=None, q=None, q_prefilled=False):
        n = p.numel()
        if q is None:
            q = self._tensor("q", n, torch.float64)
        wsb = int(self.lib.fhc_bh_workspace_bytes(n))
        ws = self._buf("bh_ws", wsb)
        # one host sync after the compaction: only the radix passes the number of ranked keys needs are launched
        ranked = ctypes.c_int64(0)
        check(self.lib.fhc_bh_qvalues_hostcount(dptr(p), n, float(T), int(rank_offset), float(carry_in), dptr(q),
                                                dptr(carry_out), dptr(n_sorted_out), ctypes.byref(ranked), dptr(ws), wsb,
                                                1 if q_prefilled else 0, self._stream()))
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
        # {flagged, first duplicate = UINT64_MAX}; from a device-resident template: no host->device copy per run
        if getattr(self, "_stats0", None) is None:
            self._stats0 = torch.tensor([0, -1], dtype=torch.int64, device=self.device)
        return outl, self._stats0.clone()

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


VARSIZE_GPU_MIN_PAIRS = 200_000_000  # -r 0: more fragment pairs in range than this go to the prefix-sum kernel


def varsize_pairs_in_range(mids, off, L, U):
    """How many fragment pairs (x < y, same chromosome) have L <= mid_y - mid_x <= U (-1: unbounded), by bisection."""
    total = 0
    for c in range(len(off) - 1):
        f = mids[off[c]:off[c + 1]]
        n = len(f)
        if n < 2:
            continue
        idx = np.arange(n, dtype=np.int64)
        lo = idx + 1 if L < 0 else np.maximum(np.searchsorted(f, f + L, side="left"), idx + 1)
        hi = np.full(n, n, dtype=np.int64) if U < 0 else np.searchsorted(f, f + U, side="right")
        total += int(np.maximum(hi - lo, 0).sum())
    return total


def frag_pairs(lib, frags, st, bins, dec=None, stream=None):
    """generate_FragPairs (fithic/fithic.py:596-689 fixed-size bins, :691-778 restriction fragments).  Mutates `bins`
    (adds pairs = `[1]`, pairs7 = `[7]`, sumdist = `[3]`; the two pair counts differ only for restriction fragments).
    Restriction fragments: the pair-by-pair walk on the host (the reference's bits in `[3]`) up to VARSIZE_GPU_MIN_PAIRS
    pairs in range, the prefix-sum kernel beyond (FHC_VARSIZE_PAIRS=host|gpu forces one)."""
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
        how = os.environ.get("FHC_VARSIZE_PAIRS", "auto")
        if how == "auto":
            how = "gpu" if varsize_pairs_in_range(mids, off, st.L, st.U) > VARSIZE_GPU_MIN_PAIRS else "host"
        lb, ub = np.ascontiguousarray(bins["lb"], dtype=np.int64), np.ascontiguousarray(bins["ub"], dtype=np.int64)
        if how == "gpu":
            check(lib.fhc_frag_pairs_varsize(dptr(mids), dptr(off), len(order), st.L, st.U, dptr(lb), dptr(ub), nb, dptr(pairs),
                                             dptr(pairs7), dptr(sumdist), dptr(totals), stream))
        else:
            check(lib.fhc_host_frag_pairs_varsize(dptr(mids), dptr(off), len(order), st.L, st.U, dptr(lb), dptr(ub), nb,
                                                  dptr(pairs), dptr(pairs7), dptr(sumdist), dptr(totals)))
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
    UnivariateSpline with s = min(y)^2 -- FITPACK's curfit restated in C (fhc_host_curfit: knots and coefficients equal
    scipy's bit for bit); the per-distance evaluation is K2."""
    y = [f for _, f in sorted(zip(x, y), key=lambda pair: pair[0])]
    x = sorted(x)
    for i in range(1, len(x)):
        if x[i] <= x[i - 1]:
            print("ERROR in spline fitting. Distances do not decrease across bins. Ensure interaction file is correct.")
            print("Avg. distance of bin(i-1)... %s" % x[i - 1])
            print("Avg. distance of bin(i)... %s" % x[i])
            raise SystemExit(2)
    m = len(x)
    if m <= 3:
        raise ValueError("the spline fit needs more than 3 bins (got %d)" % m)
    xa, ya = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    t, c = np.zeros(m + 4, dtype=np.float64), np.zeros(m + 4, dtype=np.float64)
    n, ier, calls = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
    fp = ctypes.c_double(0.0)
    check(_capi.load().fhc_host_curfit(dptr(xa), dptr(ya), m, float(min(y) * min(y)), dptr(t), dptr(c), ctypes.byref(n),
                                       ctypes.byref(fp), ctypes.byref(ier), ctypes.byref(calls)))
    return x, y, (t[:n.value].copy(), c[:n.value].copy(), 3)
