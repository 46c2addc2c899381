#!/usr/bin/env python
"""Benchmark of the Fit-Hi-C significance path on B200 (contract: see the build brief; one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--res R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`, config[3]): synthetic whole-genome intraOnly, 5 kb bins, ICE-like bias vector,
~300 M contact pairs, 1 spline pass.  A "step" is the whole path over that input: K1 histogram -> host binning + spline
fit -> K2 table -> K3 p-values -> K4 q-values.  `value` = contact pairs scored per second with the contacts resident in
HBM; `e2e` = the same through fithic_b200.api.significance with pinned HOST arrays in and out (12 B/pair H2D: mid1, mid2,
count; the chromosome ids travel run-length encoded, as the reader delivers them; D2H: p and
ExpCC whole, q as the (line, value) pairs that differ from 1.0 -- all inside the timed region).  With N GPUs the same 300 M pairs are sharded by chromosome (strong scaling).
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
    "hist_distance_kernel": ("pairs", 16, 1),      # 4 x int32 read
    "pvalues_kernel": ("pairs", 32, 1),            # tile-phased K3: 16 read + p, ExpCC written
    "pval_front_kernel": ("pairs", 32, 1),         # work-list K3, front: 16 read + p, ExpCC written (+ 16 per listed item)
    "pval_iterate_kernel": ("items", 32, 1),       # item read, numerator/denominator written
    "pval_finish_kernel": ("items", 40, 1),        # item + numerator/denominator read, p written
    "bh_cut_hist_kernel": ("pairs", 8, 1),         # p read
    "bh_compact_kernel": ("pairs", 20, 1),         # p read, (key, index) written
    "radix_upsweep_kernel": ("sorted", 8, 8),      # key read, one launch per 8-bit digit
    "radix_downsweep_kernel": ("sorted", 24, 8),   # (key, index) read and written, one launch per digit
    "bh_tilemax_kernel": ("sorted", 8, 1),
    "bh_scatter_kernel": ("sorted", 20, 1),        # (key, index) read, q written
}
PASS_BYTES_PER_PAIR = 64  # K1 16 + K3 32 + K4 16 (read p, write q)


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


def cpu_baseline(res, sample_pairs, seed, passes=1):
    from fithic_b200 import synth
    from fithic_b200.engine import Settings
    from oracle import fithic_oracle as O
    from tests.util import oracle_inputs
    cores = os.cpu_count() or 1
    O.build_c_oracle()
    key = (res, sample_pairs, seed, passes)
    if key not in _CPU_INPUTS:  # the sample is generated once per process, outside the timed part
        contacts, frags, biases, _ = synth.make_intra(sample_pairs, res, seed=seed, mean_count=3.0, with_bias=True)
        st = Settings(resolution=res, noOfBins=100, noOfPasses=passes)
        _CPU_INPUTS.clear()
        _CPU_INPUTS[key] = oracle_inputs(contacts, frags, st, biases)
    oc, fchr, fmid, fh, ost, ob = _CPU_INPUTS[key]
    t0 = time.perf_counter()
    O.run_pipeline(oc, fchr, fmid, fh, ost, ob, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": sample_pairs * passes / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d-pair whole-genome %d bp sample of the same generator, oracle/fithic_oracle.run_pipeline "
                      "(numpy + OpenMP C cephes), %.1f s" % (sample_pairs, res, dt), "seconds": dt}


def run_reference_arm(args):
    """`--impl reference`: the reference's algorithm on the host cores (oracle port; /root/reference is Python and does
    not exist on the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(args.res, max(args.ref_sample // 4, 100000), args.seed)
    for _ in range(args.steps):
        vals.append(cpu_baseline(args.res, args.ref_sample, args.seed))
    v = float(np.mean([x["value"] for x in vals]))
    ms = float(np.mean([x["seconds"] for x in vals]) * 1e3)
    cb = dict(vals[-1])
    cb["value"] = v
    cb.pop("seconds", None)
    line = {"impl": "reference", "metric": "contact-pair p-values/sec (5kb intra WG)", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "synthetic whole-genome intraOnly %d bp + ICE-like bias vector, %d contact pairs, %d spline "
                        "pass(es), sharded by chromosome" % (args.res, args.pairs, args.passes),
            "pairs": args.pairs, "resolution": args.res, "bins": 100, "passes": args.passes,
            "line_order": "sorted by (chromosome, mid1, mid2) like a contact file" if args.order == "file"
                          else "random inside each chromosome",
            "l2_policy": "inputs (16 B/pair) are far larger than the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=300_000_000)
    ap.add_argument("--res", type=int, default=5000)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1004)
    ap.add_argument("--ref-sample", type=int, default=16_000_000,
                    help="contact pairs of the CPU arm's bounded sample (16 M: ~11 s of oracle time on 16 cores)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--order", default="file", choices=["file", "random"],
                    help="line order inside a chromosome: sorted by (mid1, mid2) as contact files are, or as drawn")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from fithic_b200 import _capi, api, synth
    from fithic_b200.engine import Contacts, Engine, Settings

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
    _capi.load()

    # ---- input: this rank's chromosomes of the 300 M pair data set, generated on the device ----
    names, sizes = synth.genome(None)
    shards = synth.lpt_shards([int(s) for s in sizes], world)
    (mid1, mid2, cnt, chrs), frags, biases, per = synth.make_intra_device(
        args.pairs, args.res, args.seed, device, mean_count=3.0, with_bias=True, only=shards[rank], order=args.order)
    n_local = mid1.numel()
    st = Settings(resolution=args.res, noOfBins=100, noOfPasses=args.passes)
    eng = Engine(st, frags, biases, device=device, dist_ctx=dctx)
    eng.set_contacts_device(mid1, mid2, cnt, chrs)
    torch.cuda.synchronize()

    def step():
        outl, stats = eng.new_outlier_state()  # outlier selection runs in every pass (fithic/fithic.py:1215-1217)
        r = None
        for passNo in range(1, args.passes + 1):
            r = eng.run_pass(passNo, outl, stats)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _capi.launch_count()
    _capi.profile_enable(True)
    _capi.profile_collect()
    uuid = getattr(torch.cuda.get_device_properties(local_rank), "uuid", None)
    sampler = ClockSampler(local_rank, uuid)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(args.steps):
        last = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.pause()
    prof = _capi.profile_collect()
    _capi.profile_enable(False)
    launches = _capi.launch_count() - launches0
    # what K4 ranked: the p-values below the cut in force (bh.cu: rank bound tightened by the value histogram)
    T_last = float(last["T"])
    if world == 1:
        lib = _capi.load()
        p_cut0 = float(lib.fhc_bh_p_cut(T_last, float(n_local)))
        hist = torch.zeros(_capi.BH_CUT_BUCKETS, dtype=torch.int64, device=device)
        _capi.check(lib.fhc_bh_cut_hist(_capi.dptr(last["p"]), n_local, p_cut0, _capi.dptr(hist),
                                        ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        hh = np.ascontiguousarray(hist.cpu().numpy().view(np.uint64))
        p_cut = float(lib.fhc_host_bh_cut_find(_capi.dptr(hh), T_last, 0.0, p_cut0))
    else:
        p_cut0 = p_cut = float(dctx.last_plan["p_cut"])
        p_cut0 = float(dctx.last_plan["p_cut0"])
    n_sorted = int((last["p"] < p_cut).sum().item())
    n_below_rank_bound = int((last["p"] < p_cut0).sum().item())
    n_items = 0
    ws = eng._ws.get("pval_ws")
    if ws is not None and os.environ.get("FHC_PVAL_IMPL", "lists")[0] != "t":
        n_items = int(ws[:16].view(torch.int64).sum().item())  # continued fractions + tail sums of the last K3 launch
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = args.pairs * args.passes / (ms_per_step * 1e-3)

    # ---- e2e through the public API with pinned host buffers (H2D + D2H inside the timed region) ----
    from fithic_b200.engine import chr_runs_of
    host_chrs = chrs.cpu().pin_memory().numpy().view(np.uint32)
    # a contact file is grouped by chromosome: the reader hands the ids over in run-length form as well (io.read_contacts),
    # and then the 4 B per line of `chrs` stay on the host
    host = Contacts(*(t.cpu().pin_memory().numpy() for t in (mid1, mid2, cnt)), host_chrs, list(names),
                    None if os.environ.get("FHC_BENCH_DENSE_CHRS") else chr_runs_of(host_chrs))
    h2d_per_pair = 12 if (host.chr_runs is not None and len(host.chr_runs[0]) <= Engine.MAX_CHR_RUNS) else 16
    out = api.HostBuffers(n_local)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    api.significance(host, frags, st, biases, engine=eng, out=out)  # warm-up
    barrier()
    if rank == 0:
        sampler.resume()
    ms0 = torch.cuda.memory_stats(device)
    e0.record()
    e2e_each = []
    for _ in range(e2e_steps):
        tc = time.perf_counter()
        res = api.significance(host, frags, st, biases, engine=eng, out=out)  # returns with the results on the host
        e2e_each.append((time.perf_counter() - tc) * 1e3)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms1 = torch.cuda.memory_stats(device)
    alloc_diag = {k: int(ms1.get(k, 0) - ms0.get(k, 0)) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    alloc_diag["reserved_gb"] = ms1.get("reserved_bytes.all.current", 0) / 1e9
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = args.pairs * args.passes / (e2e_ms / e2e_steps * 1e-3)
    checksum = float(np.nansum(res[-1]["q"][:1000]))
    # bytes that crossed the link per step: p and ExpCC whole; q as (line, value) pairs where it is not 1.0, or whole
    n_ex = int(res[-1].get("q_exceptions", -1))
    q_bytes = 8 + 12 * n_ex if 0 <= n_ex <= max(n_local // 128, 1024) else 8 * n_local
    d2h_bytes = (16 * n_local + q_bytes) * args.passes
    if world > 1:
        t = torch.tensor([d2h_bytes], dtype=torch.int64, device=device)
        dist.all_reduce(t)
        d2h_bytes = int(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (device time from CUDA events recorded after every launch) ----
    peak, peak_src = peaks()
    units = {"pairs": n_local, "sorted": n_sorted, "items": n_items}
    kern = {k: v for k, v in prof.items()}
    top = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    roofline = None
    breakdown = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps}
                 for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
    if top is not None:
        # per full-size launch: the multi-GPU path also runs the sort on a 64k-key sample (negligible bytes and time),
        # so bytes and time are both taken per step and divided by the number of full-size launches
        unit, bpu, mult = ALGO_BYTES.get(top, ("pairs", 0, 1))
        mult *= args.passes
        avg_ms = kern[top]["ms"] / args.steps / mult
        achieved = bpu * units[unit] / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if top in tj and tj[top].get("pairs") == args.pairs and tj[top].get("n_gpus") == world:
                traffic = tj[top]["dram_bytes_per_launch"]
        roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bpu * units[unit], "avg_launch_ms": avg_ms,
                    "launches_per_step": mult,
                    "note": "K3 is bound by instruction issue and FP64 latency, not HBM (see DESIGN.md section 4)"
                    if top in ("pvalues_kernel", "pval_front_kernel", "pval_iterate_kernel", "pval_finish_kernel") else None,
                    "whole_step": {"algorithmic_bytes_per_pair": PASS_BYTES_PER_PAIR,
                                   "achieved_gbs": PASS_BYTES_PER_PAIR * n_local * args.passes / (ms_per_step * 1e-3) / 1e9,
                                   "frac": PASS_BYTES_PER_PAIR * n_local * args.passes / (ms_per_step * 1e-3) / 1e9 / peak}}
    cb = None
    if not args.no_cpu_baseline and world == 1:
        # in a fresh process: the oracle's OpenMP loop shares this process badly with torch's thread pools and 12 GB of
        # pinned memory (measured 110 s here against 2 s on its own)
        try:
            res_ = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1", "--res", str(args.res), "--seed", str(args.seed), "--ref-sample",
                                   str(args.ref_sample)], capture_output=True, text=True, timeout=600)
            cb = json.loads([ln for ln in res_.stdout.splitlines() if ln.startswith("{")][-1])["cpu_baseline"]
        except Exception as exc:  # the baseline is informational; never lose the GPU numbers over it
            cb = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % exc}
    line = {"metric": "contact-pair p-values/sec (5kb intra WG)", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_per_pair * args.pairs,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "q_exceptions": n_ex,
                    "ms_per_step": e2e_ms / e2e_steps, "ms_each_host_clock": e2e_each, "allocator": alloc_diag},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cb, "kernels": breakdown,
            "host_ms_per_pass": {k: v * 1e3 for k, v in eng.timings.get(1, {}).items()},
            "sorted_pairs": n_sorted, "pairs_below_rank_bound": n_below_rank_bound, "bh_p_cut": p_cut,
            "iterated_pairs": n_items, "pval_impl": os.environ.get("FHC_PVAL_IMPL", "lists"), "checksum_q": checksum}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
