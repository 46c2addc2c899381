#!/usr/bin/env python
"""Time fhc_host_stage (the host work between K1 and K3) on this machine's cores: the 5 kb whole-genome distance axis of
the bench workload, 1 ... 8 threads.  CPU only; run it on the GPU box to see what the box's host adds to a pass."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fithic_b200 import _capi, synth  # noqa: E402
from tests.test_host_stage import _stage_io  # noqa: E402


def main():
    lib = _capi.load()
    res, nbins, npairs = 5000, 100, int(os.environ.get("PAIRS", "4000000"))
    with open("/proc/cpuinfo") as f:
        models = [ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")]
    print("cpu:", models[0] if models else "?", "x", len(models), "os.cpu_count", os.cpu_count(),
          "affinity", len(os.sched_getaffinity(0)))
    contacts, frags, _, _ = synth.make_intra(npairs, res, seed=1004, mean_count=3.0, with_bias=True)
    d = np.abs(contacts.mid1.astype(np.int64) - contacts.mid2)
    D = int(max(d.max(), frags.max_mid.max()) // res + 2)
    hist = np.bincount(d // res, weights=contacts.cnt, minlength=D).astype(np.int64)
    hist[hist == 0] = 1  # the 300 M pair input observes (almost) every distance
    scal = np.zeros(_capi.N_SCALARS, dtype=np.uint64)
    scal[_capi.S_INTRA_INRANGE_SUM] = int(hist.sum())
    scal[_capi.S_MAX_COUNT] = int(contacts.cnt.max())
    for threads in (1, 2, 3, 4, 6, 8):
        io, keep = _stage_io(lib, hist, scal, None, res, nbins, frags, 0, -1, 1, threads)
        ts, parts = [], []
        for _ in range(30):
            lib.fhc_host_pool_prewarm(threads)
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 100e-6:  # K1 runs about this long at 8 GPUs
                pass
            t0 = time.perf_counter()
            _capi.check(lib.fhc_host_stage(ctypes.byref(io), 7))
            ts.append((time.perf_counter() - t0) * 1e3)
            parts.append(list(io.timings))
        parts = np.median(np.array(parts), axis=0)
        print("threads %d: stage median %.3f ms (min %.3f)  bins %.3f pairs+lbeta %.3f fit %.3f eval %.3f antitonic %.3f "
              "lut %.3f   [m = %d, knots = %d, fit calls = %d]" % (threads, np.median(ts), min(ts), parts[0], parts[1],
                                                                    parts[2], parts[3], parts[4], parts[5], io.m, io.nt,
                                                                    io.calls))
    n = ctypes.c_int32()
    for threads in (2, 4, 8):
        lib.fhc_host_pool_prewarm(threads)
        ms = [lib.fhc_host_pool_selftest(threads, 32, 25, ctypes.byref(n)) for _ in range(5)]
        print("pool selftest %d threads: 32 jobs x 25 us -> %s ms (ideal %.3f)" % (threads, ["%.3f" % v for v in ms],
                                                                               32 * 25e-3 / threads))


if __name__ == "__main__":
    main()
