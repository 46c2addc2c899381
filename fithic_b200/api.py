"""Public array-level API: host contact arrays in, host p / q / ExpCC out.

    from fithic_b200 import api
    passes = api.significance(contacts, fragments, settings, biases=None)

This is the call a user (or the `fithic` CLI) makes; bench.py times it end to end ("e2e"): every call copies the
contact arrays host -> device and the three result arrays device -> host.
"""
import os
import threading

import numpy as np
import torch

from . import _capi

from .engine import Biases, Contacts, Engine, Fragments, Settings  # noqa: F401  (re-exported)


def _fill_threads():
    """Host threads for the q fill: the cores this rank can call its own (all of them on one GPU, a share under torchrun)."""
    env = os.environ.get("FHC_FILL_THREADS")
    if env:
        return max(1, int(env))
    cores = os.cpu_count() or 4
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(2, min(16, cores // max(local, 1)))


class HostBuffers:
    """Reusable pinned host buffers for the results of one run (3 x 8 B per contact)."""

    def __init__(self, n):
        self.n = n
        self.p = torch.empty(n, dtype=torch.float64).pin_memory()
        self.q = torch.empty(n, dtype=torch.float64).pin_memory()
        self.expcc = torch.empty(n, dtype=torch.float64).pin_memory()
        # what significance() knows about q from its previous call with these buffers: None = nothing (fill it), else the
        # lines that differ from 1.0 (only those have to be reset).  A caller that writes into q calls invalidate().
        self.q_not_one = None

    def invalidate(self):
        self.q_not_one = None


def significance(contacts, fragments, settings, biases=None, engine=None, out=None):
    """Run every spline pass on the current CUDA device.

    contacts: engine.Contacts (numpy int32/uint32 arrays, ideally views of pinned memory).
    Returns a list with one dict per pass: p, q, expcc (numpy float64, file order) plus the per-pass tables
    (bins, x, y, spline knots, N, T, outlier threshold).  `out` (HostBuffers) is reused for the last pass's arrays.
    """
    eng = engine if engine is not None else Engine(settings, fragments, biases)
    eng.upload_contacts(contacts, non_blocking=True)
    n = len(contacts)
    if out is None or out.n != n:
        out = HostBuffers(n)
    outl, stats = eng.new_outlier_state()
    main = torch.cuda.current_stream(eng.device)
    copy = getattr(eng, "_copy_stream", None)
    if copy is None:
        copy = eng._copy_stream = torch.cuda.Stream(device=eng.device)
    results = []
    for passNo in range(1, settings.noOfPasses + 1):
        if passNo > 1 and settings.interOnly:
            break

        def copy_slice(lo, hi, p, e):
            # finished slices of p and ExpCC leave for the host on a second stream while K3 works on the next slice and
            # K4 (sort + scan) runs afterwards: the PCIe link is the bottleneck of the end-to-end call
            copy.wait_stream(main)
            with torch.cuda.stream(copy):
                out.p[lo:hi].copy_(p[lo:hi], non_blocking=True)
                out.expcc[lo:hi].copy_(e[lo:hi], non_blocking=True)

        # q is 1.0 on almost every line of a sparse map: a host thread fills the pinned array with 1.0 while the GPU works
        # and only the (line, q) pairs that differ cross the PCIe link (dense copy when they are more than n / 128)
        if out.q_not_one is None:
            filler = threading.Thread(target=eng.lib.fhc_host_fill_f64, args=(_capi.dptr(out.q), n, 1.0, _fill_threads()))
        else:  # the buffers come from an earlier call: everything is 1.0 already except the lines recorded then
            filler = threading.Thread(target=out.q.numpy().__setitem__, args=(out.q_not_one, 1.0))
        filler.start()
        r = eng.run_pass(passNo, outl, stats, pvalue_chunks=8 if n >= (1 << 22) else 1, after_chunk=copy_slice)
        cap = max(n // 128, 1024)
        ex_idx = eng._tensor("q_ex_idx", cap, torch.int32)
        ex_val = eng._tensor("q_ex_val", cap, torch.float64)
        ex_cnt = eng._tensor("q_ex_cnt", 1, torch.int64)
        _capi.check(eng.lib.fhc_gather_ne_one(_capi.dptr(r["q"]), n, cap, _capi.dptr(ex_idx), _capi.dptr(ex_val),
                                              _capi.dptr(ex_cnt), eng._stream()))
        n_ex = int(ex_cnt.item())
        filler.join()
        if n_ex <= cap:
            lines = ex_idx[:n_ex].cpu().numpy().view(np.uint32).astype(np.int64)
            if n_ex:
                out.q.numpy()[lines] = ex_val[:n_ex].cpu().numpy()
            out.q_not_one = lines
        else:
            out.q.copy_(r["q"], non_blocking=True)
            out.q_not_one = None
        r["q_exceptions"] = n_ex
        main.synchronize()
        copy.synchronize()
        last = passNo == settings.noOfPasses or settings.interOnly
        r["p"] = out.p.numpy() if last else out.p.numpy().copy()
        r["q"] = out.q.numpy() if last else out.q.numpy().copy()
        r["expcc"] = out.expcc.numpy() if last else out.expcc.numpy().copy()
        results.append(r)
    return results
