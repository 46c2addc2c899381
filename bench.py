#!/usr/bin/env python
"""Benchmark of the Fit-Hi-C significance path on B200 (contract: see the build brief; one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c4|c2|c3|c5] [--extras LIST]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Main line (BASELINE.json `metric`, configs[3], "c4"): synthetic whole-genome intraOnly, 5 kb bins, ICE-like bias vector,
~300 M contact pairs, 1 spline pass.  A "step" is the whole path over that input: K1 histogram -> host binning + spline
fit -> K2 table -> K3 p-values -> K4 q-values.  `value` = contact pairs scored per second with the contacts resident in
HBM; `e2e` = the same through fithic_b200.api.significance with pinned HOST arrays in and out (12 B/pair H2D: mid1, mid2,
count; the chromosome ids travel run-length encoded, as the reader delivers them; D2H: p and ExpCC whole, q as the (line,
value) pairs that differ from 1.0 -- all inside the timed region).  With N GPUs the same pairs are sharded by chromosome
(strong scaling; the workload cannot grow with N because the sum of the counts must stay below 2^31).

`extra` carries short device-resident runs of the other BASELINE.json configs (c2: chr1 40 kb 2 M pairs; c3: whole genome
10 kb 80 M pairs 2 passes; c5: whole genome 25 kb interOnly 100 M lines) and of the main workload with planted signal (8 %
of the lines get extra reads, so that a share of the q-values like on real maps falls below 1 and K4 has to rank them).
`digest_line_p_q` is an order-independent hash of (file line, p, q) over the whole file: equal at every N iff every line
got the same p and q.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {  # kernel: (unit, algorithmic HBM bytes per unit per launch, full-size launches per pass); DESIGN.md 4
    "hist_distance_kernel": ("pairs", 12, 1),      # 3 x int32 read (the chromosome ids come as runs)
    "pvalues_kernel": ("pairs", 32, 1),            # tile-phased K3: 16 read + p, ExpCC written
    "pval_front_kernel": ("pairs", 28, 1),         # work-list K3, front: 12 read + p, ExpCC written (+ 16 per listed item)
    "pval_front2_kernel": ("pairs", 28, 1),        # its second version (the default); the first one runs under FHC_PVAL_FRONT=v1
    "pval_iterate_kernel": ("items", 32, 1),       # item read, numerator/denominator written
    "pval_finish_kernel": ("items", 40, 1),        # item + numerator/denominator read, p written
    "bh_cut_hist_kernel": ("pairs", 8, 1),         # p read
    "bh_compact_kernel": ("pairs", 20, 1),         # p read, (key, index) written
    "bh_part_scatter_kernel": ("pairs", 16, 1),    # multi-GPU: p read, q written
    "radix_upsweep_kernel": ("sorted", 8, 8),      # key read, one launch per 8-bit digit
    "radix_downsweep_kernel": ("sorted", 24, 8),   # (key, index) read and written, one launch per digit
    "bh_tilemax_kernel": ("sorted", 8, 1),
    "bh_scatter_kernel": ("sorted", 20, 1),        # (key, index) read, q written
}
PASS_BYTES_PER_PAIR = 56  # K1 12 + K3 28 + K4 16 (read p, write q)
K3_KERNELS = ("pvalues_kernel", "pval_front_kernel", "pval_front2_kernel", "pval_iterate_kernel", "pval_finish_kernel")

CONFIGS = {
    # BASELINE.json configs[3]: the configuration the metric is quoted on
    "c4": dict(kind="intra", res=5000, pairs=300_000_000, bias=True, passes=1, mean_count=3.0, seed=1004, chroms=None,
               label="synthetic whole-genome intraOnly 5000 bp + ICE-like bias vector"),
    "c2": dict(kind="intra", res=40000, pairs=2_000_000, bias=False, passes=1, mean_count=8.0, seed=1002, chroms=["chr1"],
               label="synthetic chr1 intraOnly 40000 bp, no bias"),
    "c3": dict(kind="intra", res=10000, pairs=80_000_000, bias=False, passes=2, mean_count=4.0, seed=1003, chroms=None,
               label="synthetic whole-genome intraOnly 10000 bp, no bias"),
    "c5": dict(kind="inter", res=25000, pairs=100_000_000, bias=False, passes=1, seed=1005, intra_fraction=0.1,
               label="synthetic whole-genome interOnly 25000 bp (10 % intra lines), constant prior + global BH"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: NVML polled every few ms from a thread (the device-
    resident leg lasts ~0.1 s, too short for `nvidia-smi -lms`); falls back to one nvidia-smi query per call if NVML is
    unavailable."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0, uuid=None, period=0.004):
        self.index, self.uuid, self.period = index, uuid, period
        self.samples = []   # (sm_mhz, reasons bitmask, power W)
        self.sm_max = None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid is not None:
                try:
                    u = str(uuid)
                    h = pynvml.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = index
                if vis:
                    try:
                        phys = int(vis.split(",")[index])
                    except Exception:
                        phys = index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv, self._h = pynvml, h
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _poll(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    try:
                        rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    try:
                        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                    except Exception:
                        pw = None
                    self.samples.append((sm, rs, pw))
                except Exception:
                    pass
            time.sleep(self.period)

    def start(self):
        if self._h is not None and self._thread is None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        self.resume()

    def resume(self):
        self._active.set()
        if self._h is None:
            self._smi_sample()

    def pause(self):
        self._active.clear()

    def _smi_sample(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            rs = 0
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    rs |= self.BAD[name]
            self.samples.append((float(f[0]), rs, float(f[2])))
            self.sm_max = float(f[1])
        except Exception:
            pass

    def stop(self):
        self._active.clear()
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"], "samples": 0}
        sm = [x[0] for x in self.samples]
        reasons = sorted(n for n, bit in self.BAD.items() if any(x[1] & bit for x in self.samples))
        pw = [x[2] for x in self.samples if x[2] is not None]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "source": "nvml" if self._h is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (numpy restatement + OpenMP C cephes) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
_CPU_INPUTS = {}


def cpu_baseline(cfg, sample_pairs):
    from fithic_b200 import synth
    from fithic_b200.engine import Settings
    from oracle import fithic_oracle as O
    from tests.util import oracle_inputs
    cores = os.cpu_count() or 1
    O.build_c_oracle()
    res, seed, passes = cfg["res"], cfg["seed"], cfg["passes"]
    key = (cfg["label"], sample_pairs)
    if key not in _CPU_INPUTS:  # the sample is generated once per process, outside the timed part
        if cfg["kind"] == "inter":
            contacts, frags, biases, _ = synth.make_intra(sample_pairs // 10, res, seed=seed, mean_count=1.3, with_bias=False,
                                                          inter_fraction=9.0)
            st = Settings(resolution=res, noOfBins=100, interOnly=True)
        else:
            contacts, frags, biases, _ = synth.make_intra(sample_pairs, res, seed=seed, mean_count=cfg["mean_count"],
                                                          with_bias=cfg["bias"], chroms=cfg["chroms"])
            st = Settings(resolution=res, noOfBins=100, noOfPasses=passes)
        _CPU_INPUTS.clear()
        _CPU_INPUTS[key] = (oracle_inputs(contacts, frags, st, biases), len(contacts))
    (oc, fchr, fmid, fh, ost, ob), nlines = _CPU_INPUTS[key]
    t0 = time.perf_counter()
    O.run_pipeline(oc, fchr, fmid, fh, ost, ob, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": nlines * passes / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d-line sample of the same generator (%s), oracle/fithic_oracle.run_pipeline (numpy + OpenMP C cephes), "
                      "%d pass(es), %.1f s; the unmodified single-threaded reference runs at ~2.5e4 pairs/s (BASELINE.md)"
                      % (nlines, cfg["label"], passes, dt), "seconds": dt}


def run_reference_arm(args, cfg):
    """`--impl reference`: the reference's algorithm on the host cores (oracle port; /root/reference is Python and does
    not exist on the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    sample = min(args.ref_sample, cfg["pairs"])
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(cfg, max(sample // 4, 100000))
    for _ in range(args.steps):
        vals.append(cpu_baseline(cfg, sample))
    v = float(np.mean([x["value"] for x in vals]))
    ms = float(np.mean([x["seconds"] for x in vals]) * 1e3)
    cb = dict(vals[-1])
    cb["value"] = v
    cb.pop("seconds", None)
    line = {"impl": "reference", "metric": "contact-pair p-values/sec (5kb intra WG)", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, cfg, signal=0.0):
    return {"workload": "%s, %d contact pairs, %d spline pass(es), sharded by chromosome%s"
                        % (cfg["label"], cfg["pairs"], cfg["passes"],
                           ", %.0f %% of the lines with planted extra reads" % (100 * signal) if signal else ""),
            "name": args.config, "pairs": cfg["pairs"], "resolution": cfg["res"], "bins": 100, "passes": cfg["passes"],
            "line_order": "sorted by (chromosome, mid1, mid2) like a contact file" if args.order == "file"
                          else "random inside each chromosome",
            "l2_policy": "inputs (12 B/pair) are far larger than the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """One configuration on this rank: device-resident contacts, engine, the step function."""

    def __init__(self, cfg, args, device, dctx, world, rank, signal=0.0):
        import torch
        from fithic_b200 import synth
        from fithic_b200.engine import Engine, Settings
        self.cfg, self.signal = cfg, signal
        self.world, self.rank, self.device, self.dctx = world, rank, device, dctx
        res = cfg["res"]
        self.passes = cfg["passes"]
        if cfg["kind"] == "inter":
            chunk = 1 << 22
            nchunks = (cfg["pairs"] + chunk - 1) // chunk
            lo, hi = (nchunks * rank) // world, (nchunks * (rank + 1)) // world
            (mid1, mid2, cnt, chrs), frags, first = synth.make_inter_device(cfg["pairs"], res, cfg["seed"], device,
                                                                           cfg["intra_fraction"], chunk, range(lo, hi))
            biases = None
            self.names = synth.genome(None)[0]
            st = Settings(resolution=res, noOfBins=100, interOnly=True)
            eng = Engine(st, frags, biases, device=device, dist_ctx=dctx)
            eng.set_contacts_device(mid1, mid2, cnt, chrs)
            eng.set_line_runs([first], [mid1.numel()])
            self.host_runs = None
        else:
            names, sizes = synth.genome(cfg["chroms"])
            shards = synth.lpt_shards([int(s) for s in sizes], world)
            (mid1, mid2, cnt, chrs), frags, biases, per = synth.make_intra_device(
                cfg["pairs"], res, cfg["seed"], device, mean_count=cfg["mean_count"], with_bias=cfg["bias"],
                only=shards[rank], order=args.order, chroms=cfg["chroms"], signal_frac=signal)
            self.names = names
            st = Settings(resolution=res, noOfBins=100, noOfPasses=self.passes)
            eng = Engine(st, frags, biases, device=device, dist_ctx=dctx)
            # the lines of a chromosome are consecutive in the file (and in this rank's shard): chromosome ids and file
            # positions as runs
            mine = [c for c in shards[rank] if per[c] > 0]
            run_vals = np.array([c | (c << 16) for c in mine], dtype=np.uint32)
            run_lens = np.array([per[c] for c in mine], dtype=np.int64)
            file_start = np.concatenate([[0], np.cumsum(per)[:-1]]).astype(np.int64)
            self.host_runs = (run_vals, run_lens)
            eng.set_contacts_device(mid1, mid2, cnt, chrs, chr_runs=self.host_runs)
            eng.set_line_runs(file_start[mine], run_lens)
        self.contacts = (mid1, mid2, cnt, chrs)
        self.frags, self.biases, self.st, self.eng = frags, biases, st, eng
        self.n_local = mid1.numel()
        torch.cuda.synchronize()

    def step(self):
        eng = self.eng
        outl, stats = eng.new_outlier_state()  # outlier selection runs in every pass (fithic/fithic.py:1215-1217)
        r = None
        for passNo in range(1, self.passes + 1):
            r = eng.run_pass(passNo, outl, stats)
        return r

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, warmup, sampler=None):
        """W untimed steps, then exactly K steps between barriers; device time from CUDA events, max over ranks."""
        import torch
        import torch.distributed as dist
        from fithic_b200 import _capi
        for _ in range(warmup):
            self.step()
        self.barrier()
        launches0 = _capi.launch_count()
        _capi.profile_enable(True)
        _capi.profile_collect()
        self.barrier()
        if sampler is not None:
            sampler.resume()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = self.step()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if sampler is not None:
            sampler.pause()
        prof = _capi.profile_collect()
        _capi.profile_enable(False)
        launches = _capi.launch_count() - launches0
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return last, ms / steps, prof, launches

    def digest(self, last):
        """(file line, p, q) digest of the whole file, lines with q < 1 and lines K4 ranked, summed over the ranks."""
        import torch
        import torch.distributed as dist
        dg = np.array(self.eng.digest(last["p"], last["q"]), dtype=np.uint64)
        stats = np.array([int((last["q"] < 1.0).sum().item()), 0], dtype=np.int64)
        if self.world > 1:
            t = torch.from_numpy(np.concatenate([dg.view(np.int64), stats])).to(self.device)
            dist.all_reduce(t)  # sums wrap mod 2^64
            t = t.cpu().numpy()
            dg, stats = t[:2].copy().view(np.uint64), t[2:]
        return "%016x%016x" % (int(dg[0]), int(dg[1])), int(stats[0])


def kernel_breakdown(prof, steps):
    return {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}


def ranked_counts(w, last):
    """What K4 ranked on this rank in the last pass: p-values below the cut in force, and below the rank bound alone."""
    import torch
    from fithic_b200 import _capi
    T_last = float(last["T"])
    if w.world == 1:
        lib = _capi.load()
        p_cut0 = float(lib.fhc_bh_p_cut(T_last, float(w.n_local)))
        hist = torch.zeros(_capi.BH_CUT_BUCKETS, dtype=torch.int64, device=w.device)
        _capi.check(lib.fhc_bh_cut_hist(_capi.dptr(last["p"]), w.n_local, p_cut0, _capi.dptr(hist),
                                        ctypes.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)))
        hh = np.ascontiguousarray(hist.cpu().numpy().view(np.uint64))
        p_cut = float(lib.fhc_host_bh_cut_find(_capi.dptr(hh), T_last, 0.0, p_cut0))
    else:
        p_cut = float(w.dctx.last_plan["p_cut"])
        p_cut0 = float(w.dctx.last_plan["p_cut0"])
    return int((last["p"] < p_cut).sum().item()), int((last["p"] < p_cut0).sum().item()), p_cut


def short_run(name, cfg, args, device, dctx, world, rank, signal=0.0, steps=5, warmup=3):
    """A device-resident timing of one more configuration (no end-to-end leg): one entry of `extra`."""
    w = Workload(cfg, args, device, dctx, world, rank, signal)
    last, ms_per_step, prof, launches = w.timed(steps, warmup)
    digest, q_below_one = w.digest(last)
    n_sorted, n_rank_bound, p_cut = ranked_counts(w, last)
    out = {"name": name, "workload": workload_config(args, cfg, signal)["workload"], "pairs": cfg["pairs"],
           "passes": cfg["passes"], "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
           "value": cfg["pairs"] * cfg["passes"] / (ms_per_step * 1e-3), "unit": "pairs/s",
           "lines_with_q_below_1": q_below_one, "sorted_pairs_rank0": n_sorted, "bh_p_cut": p_cut,
           "digest_line_p_q": digest, "kernels": kernel_breakdown(prof, steps),
           "host_ms_per_pass": {k: v * 1e3 for k, v in w.eng.timings.get(cfg["passes"], {}).items()}}
    k4 = sum(v["ms_per_step"] for k, v in out["kernels"].items() if k.startswith(("bh_", "radix_", "scan_", "sort_")))
    out["k4_ms_per_step"] = k4
    del w
    import torch
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS), help="BASELINE.json configuration of the main line")
    ap.add_argument("--pairs", type=int, default=None, help="override the configuration's number of contact pairs")
    ap.add_argument("--passes", type=int, default=None)
    ap.add_argument("--signal", type=float, default=0.0, help="share of the lines with planted extra reads (main line)")
    ap.add_argument("--extras", default="c2,c3,c5,signal",
                    help="comma list of short extra runs reported under `extra` (c2, c3, c5, signal; '' = none)")
    ap.add_argument("--ref-sample", type=int, default=16_000_000,
                    help="contact pairs of the CPU arm's bounded sample (16 M: ~11 s of oracle time on 16 cores)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--order", default="file", choices=["file", "random"],
                    help="line order inside a chromosome: sorted by (mid1, mid2) as contact files are, or as drawn")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.pairs:
        cfg["pairs"] = args.pairs
    if args.passes:
        cfg["passes"] = args.passes
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    args.warmup = max(args.warmup, 3)  # timing rules: at least three untimed steps

    import torch
    import torch.distributed as dist
    from fithic_b200 import _capi, api
    from fithic_b200.engine import Contacts, Engine, host_threads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dctx = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL may print its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        from fithic_b200.parallel import DistCtx
        dctx = DistCtx(device)
    lib = _capi.load()

    # ---- FP64 peak of this GPU (the roofline K3 is reported against) ----
    fp64_peak = None
    if rank == 0:
        scratch = torch.zeros(1, dtype=torch.float64, device=device)
        tf = ctypes.c_double(0.0)
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        if lib.fhc_peak_fp64(0.0, _capi.dptr(scratch), ctypes.byref(tf), stream) == 0:
            burst = tf.value
            _capi.check(lib.fhc_peak_fp64(0.5, _capi.dptr(scratch), ctypes.byref(tf), stream))
            fp64_peak = {"burst_tflops": burst, "sustained_tflops": tf.value,
                         "how": "fhc_peak_fp64: 16 independent DFMA chains per thread, 148 x 8 CTAs of 256 threads; best of "
                                "five 10 ms bursts / back to back for 0.5 s"}

    # ---- main workload: this rank's chromosomes of the data set, generated on the device ----
    w = Workload(cfg, args, device, dctx, world, rank, args.signal)
    eng, st, n_local = w.eng, w.st, w.n_local
    uuid = getattr(torch.cuda.get_device_properties(local_rank), "uuid", None)
    sampler = ClockSampler(local_rank, uuid)
    if rank == 0:
        sampler.start()
        sampler.pause()
    last, ms_per_step, prof, launches = w.timed(args.steps, args.warmup, sampler if rank == 0 else None)
    digest, q_below_one = w.digest(last)
    n_sorted, n_below_rank_bound, p_cut = ranked_counts(w, last)
    n_items = 0
    ws = eng._ws.get("pval_ws")
    if ws is not None and os.environ.get("FHC_PVAL_IMPL", "lists")[0] != "t":
        packed = int(ws[:8].view(torch.int64).item())  # list lengths of the last K3 launch in one word (pvalue_lists.cu):
        n_items = (packed & 0xffffffff) + ((packed >> 32) & 0xffffffff)  # continued fractions (low) + tail sums (high)
    value = cfg["pairs"] * cfg["passes"] / (ms_per_step * 1e-3)
    host_ms = {k: v * 1e3 for k, v in eng.timings.get(cfg["passes"], {}).items()}

    # ---- e2e through the public API with pinned host buffers (H2D + D2H inside the timed region) ----
    mid1, mid2, cnt, chrs = w.contacts
    host_chrs = chrs.cpu().pin_memory().numpy().view(np.uint32)
    # a contact file is grouped by chromosome: the reader hands the ids over in run-length form as well (io.read_contacts),
    # and then the 4 B per line of `chrs` stay on the host
    host = Contacts(*(t.cpu().pin_memory().numpy() for t in (mid1, mid2, cnt)), host_chrs, list(w.names),
                    None if os.environ.get("FHC_BENCH_DENSE_CHRS") else w.host_runs)
    h2d_per_pair = 12 if (host.chr_runs is not None and len(host.chr_runs[0]) <= Engine.MAX_CHR_RUNS) else 16
    out = api.HostBuffers(n_local)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    api.significance(host, w.frags, st, w.biases, engine=eng, out=out)  # warm-up
    w.barrier()
    if rank == 0:
        sampler.resume()
    ms0 = torch.cuda.memory_stats(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_each = []
    res = None
    for _ in range(e2e_steps):
        tc = time.perf_counter()
        res = api.significance(host, w.frags, st, w.biases, engine=eng, out=out)  # returns with the results on the host
        e2e_each.append((time.perf_counter() - tc) * 1e3)
    e1.record()
    w.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms1 = torch.cuda.memory_stats(device)
    alloc_diag = {k: int(ms1.get(k, 0) - ms0.get(k, 0)) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    alloc_diag["reserved_gb"] = ms1.get("reserved_bytes.all.current", 0) / 1e9
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = cfg["pairs"] * cfg["passes"] / (e2e_ms / e2e_steps * 1e-3)
    checksum = float(np.nansum(res[-1]["q"][:1000]))
    # bytes that crossed the link per step: p and ExpCC whole; q as (line, value) pairs where it is not 1.0, or whole
    n_ex = int(res[-1].get("q_exceptions", -1))
    q_bytes = 8 + 12 * n_ex if 0 <= n_ex <= max(n_local // 128, 1024) else 8 * n_local
    d2h_bytes = (16 * n_local + q_bytes) * cfg["passes"]
    if world > 1:
        t = torch.tensor([d2h_bytes], dtype=torch.int64, device=device)
        dist.all_reduce(t)
        d2h_bytes = int(t.item())
    e2e_host = {k: round(v, 3) for k, v in getattr(api, "LAST_TIMELINE", {}).items()}
    del host, out, res
    mid1 = mid2 = cnt = chrs = None

    # ---- the other configurations and the signal workload, device resident, a few steps each ----
    extras = []
    main_kernels = kernel_breakdown(prof, args.steps)
    del w, eng, last
    torch.cuda.empty_cache()
    for name in [x for x in args.extras.split(",") if x]:
        try:
            if name == "signal":
                extras.append(short_run("c4+signal" if args.config == "c4" else args.config + "+signal", cfg, args, device,
                                        dctx, world, rank, signal=0.08))
            elif name in CONFIGS and name != args.config:
                extras.append(short_run(name, dict(CONFIGS[name]), args, device, dctx, world, rank))
        except Exception as exc:  # an extra must never cost the main line
            extras.append({"name": name, "failed": repr(exc)})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (device time from CUDA events recorded after every launch) ----
    peak, peak_src = peaks()
    units = {"pairs": n_local, "sorted": n_sorted, "items": n_items}
    kern = prof
    top = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    roofline = None
    if top is not None:
        # per full-size launch: the multi-GPU path also runs the sort on a 64k-key sample (negligible bytes and time),
        # so bytes and time are both taken per step and divided by the number of full-size launches
        unit, bpu, mult = ALGO_BYTES.get(top, ("pairs", 0, 1))
        mult *= cfg["passes"]
        avg_ms = kern[top]["ms"] / args.steps / mult
        achieved = bpu * units[unit] / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        tj = {}
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if top in tj and tj[top].get("pairs") == cfg["pairs"] and tj[top].get("n_gpus") == world:
                traffic = tj[top]["dram_bytes_per_launch"]
        step_bytes = PASS_BYTES_PER_PAIR * n_local * cfg["passes"]
        roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bpu * units[unit], "avg_launch_ms": avg_ms,
                    "launches_per_step": mult,
                    "note": "K3 is bound by instruction issue and FP64 latency, not HBM: see `fp64`" if top in K3_KERNELS
                    else None,
                    "whole_step": {"algorithmic_bytes_per_pair": PASS_BYTES_PER_PAIR,
                                   "achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                                   "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}
        # the compute side of K3: FP64 operations per unit from an ncu pass (profiles/fp64_ops.json: thread-level
        # dfma / dmul / dadd counts per kernel on this workload) against the measured FP64 FMA rate of this GPU
        fpath = os.path.join(ROOT, "profiles", "fp64_ops.json")
        if fp64_peak is not None and os.path.exists(fpath):
            with open(fpath) as f:
                fo = json.load(f)
            fp = {}
            for k in K3_KERNELS:
                if k in kern and k in fo and kern[k]["ms"] > 0:
                    unit_k = ALGO_BYTES[k][0]
                    flops = fo[k]["flops_per_unit"] * units[unit_k] * cfg["passes"]
                    tfl = flops / (kern[k]["ms"] / args.steps * 1e-3) / 1e12
                    fp[k] = {"flops_per_" + unit_k[:-1]: fo[k]["flops_per_unit"], "achieved_tflops": tfl,
                             "frac_of_burst_peak": tfl / fp64_peak["burst_tflops"]}
            roofline["fp64"] = {"peak": fp64_peak, "kernels": fp,
                                "source": "profiles/fp64_ops.json (ncu smsp__sass_thread_inst_executed_op_d*_pred_on, "
                                          "2 flops per dfma)"}
        elif fp64_peak is not None:
            roofline["fp64"] = {"peak": fp64_peak, "kernels": None}
    cb = None
    if not args.no_cpu_baseline and world == 1:
        # in a fresh process: the oracle's OpenMP loop shares this process badly with torch's thread pools and 12 GB of
        # pinned memory (measured 110 s here against 2 s on its own)
        try:
            res_ = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1", "--config", args.config, "--ref-sample", str(args.ref_sample)],
                                  capture_output=True, text=True, timeout=600)
            cb = json.loads([ln for ln in res_.stdout.splitlines() if ln.startswith("{")][-1])["cpu_baseline"]
        except Exception as exc:  # the baseline is informational; never lose the GPU numbers over it
            cb = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % exc}
    line = {"metric": "contact-pair p-values/sec (5kb intra WG)", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, args.signal), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_per_pair * cfg["pairs"],
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "q_exceptions": n_ex,
                    "ms_per_step": e2e_ms / e2e_steps, "ms_each_host_clock": e2e_each, "allocator": alloc_diag,
                    "host_timeline_ms_rank0": e2e_host},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cb, "kernels": main_kernels,
            "host_ms_per_pass": host_ms, "sorted_pairs": n_sorted, "pairs_below_rank_bound": n_below_rank_bound,
            "lines_with_q_below_1": q_below_one, "bh_p_cut": p_cut, "iterated_pairs": n_items,
            "pval_impl": os.environ.get("FHC_PVAL_IMPL", "lists"), "checksum_q": checksum, "digest_line_p_q": digest,
            "host_threads": host_threads(), "extra": extras}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
