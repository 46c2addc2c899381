#!/usr/bin/env python
"""Benchmark of the Fit-Hi-C significance path on B200 (contract: see the build brief; one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--res R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`, config[3]): synthetic whole-genome intraOnly, 5 kb bins, ICE-like bias vector,
~300 M contact pairs, 1 spline pass.  A "step" is the whole path over that input: K1 histogram -> host binning + spline
fit -> K2 table -> K3 p-values -> K4 q-values.  `value` = contact pairs scored per second with the contacts resident in
HBM; `e2e` = the same through fithic_b200.api.significance with pinned HOST arrays in and out (16 B/pair H2D, 24 B/pair
D2H inside the timed region).  With N GPUs the same 300 M pairs are sharded by chromosome (strong scaling).
"""
import argparse
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
    "pvalues_kernel": ("pairs", 32, 1),            # 16 read + p, ExpCC written
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
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (numpy restatement + OpenMP C cephes) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(res, sample_pairs, seed, passes=1):
    from fithic_b200 import synth
    from fithic_b200.engine import Settings
    from oracle import fithic_oracle as O
    from tests.util import oracle_inputs
    cores = os.cpu_count() or 1
    O.build_c_oracle()
    contacts, frags, biases, _ = synth.make_intra(sample_pairs, res, seed=seed, mean_count=3.0, with_bias=True)
    st = Settings(resolution=res, noOfBins=100, noOfPasses=passes)
    oc, fchr, fmid, fh, ost, ob = oracle_inputs(contacts, frags, st, biases)
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
    ap.add_argument("--ref-sample", type=int, default=3_000_000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        args.pairs, args.res, args.seed, device, mean_count=3.0, with_bias=True, only=shards[rank])
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
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(args.steps):
        last = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = _capi.profile_collect()
    _capi.profile_enable(False)
    launches = _capi.launch_count() - launches0
    n_sorted = int((last["p"] < 1).sum().item())
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = args.pairs * args.passes / (ms_per_step * 1e-3)

    # ---- e2e through the public API with pinned host buffers (H2D + D2H inside the timed region) ----
    host = Contacts(*(t.cpu().pin_memory().numpy() for t in (mid1, mid2, cnt)),
                    chrs.cpu().pin_memory().numpy().view(np.uint32), list(names))
    out = api.HostBuffers(n_local)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    api.significance(host, frags, st, biases, engine=eng, out=out)  # warm-up
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        res = api.significance(host, frags, st, biases, engine=eng, out=out)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = args.pairs * args.passes / (e2e_ms / e2e_steps * 1e-3)
    checksum = float(np.nansum(res[-1]["q"][:1000]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (device time from CUDA events recorded after every launch) ----
    peak, peak_src = peaks()
    units = {"pairs": n_local, "sorted": n_sorted}
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
                    "note": "pvalues_kernel is bound by instruction issue and FP64 latency, not HBM (ncu: issue slots 53 %, FP64 "
                            "pipe 20 %, DRAM 8 %; traffic = algorithmic bytes); see DESIGN.md section 4"
                    if top == "pvalues_kernel" else None,
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
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 16 * args.pairs,
                    "d2h_bytes_per_step": 24 * args.pairs * args.passes, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cb, "kernels": breakdown,
            "host_ms_per_pass": {k: v * 1e3 for k, v in eng.timings.get(1, {}).items()},
            "sorted_pairs": n_sorted, "checksum_q": checksum}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
