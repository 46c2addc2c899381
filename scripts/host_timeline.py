"""Host-side timeline of one spline pass (wall-clock between synchronisation points), for tuning fixed costs.
Run alone or under torchrun; prints per-stage milliseconds averaged over a few passes on rank 0: first the pass as the
bench runs it (no extra synchronisation), then with a device sync around every stage."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fithic_b200 import engine as E  # noqa: E402
from fithic_b200 import synth  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ctx = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    from fithic_b200.parallel import DistCtx
    ctx = DistCtx(dev)
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000_000
names, sizes = synth.genome(None)
shards = synth.lpt_shards([int(s) for s in sizes], world)
(m1, m2, c, ch), frags, biases, per = synth.make_intra_device(pairs, 5000, 1004, dev, only=shards[rank])
st = E.Settings(resolution=5000, noOfBins=100)
eng = E.Engine(st, frags, biases, device=dev, dist_ctx=ctx)
mine = [k for k in shards[rank] if per[k] > 0]
eng.set_contacts_device(m1, m2, c, ch, chr_runs=(np.array([k | (k << 16) for k in mine], dtype=np.uint32),
                                                 np.array([per[k] for k in mine], dtype=np.int64)))


def one_pass():
    o, s = eng.new_outlier_state()
    return eng.run_pass(1, o, s)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    one_pass()
barrier()
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    one_pass()
barrier()
free = (time.perf_counter() - t0) / reps * 1e3
# host time of run_pass alone (how long the host needs to enqueue a pass, GPU waits included where the host blocks)
ts = []
for _ in range(reps):
    barrier()
    t = time.perf_counter()
    one_pass()
    ts.append((time.perf_counter() - t) * 1e3)
    barrier()
stage = {k: v * 1e3 for k, v in eng.timings.get(1, {}).items()}

marks = {}


def wrap(obj, name, label):
    f = getattr(obj, name)

    def g(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize()
        marks[label] = marks.get(label, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, g)


wrap(eng, "hist_distance", "K1")
wrap(eng, "_tables_native", "D2H + host stage + H2D (fhc_host_stage)")
wrap(eng, "pvalues", "K3")
wrap(eng, "bh_qvalues", "K4 local")
if ctx is not None:
    wrap(ctx, "allreduce_k1", "exchange 1: all-reduce [hist | totals | slots]")
    wrap(ctx, "global_bh", "exchange 2 + K4 (global_bh)")
    for nme in ("cut_hist", "cut_from_hists", "partition_scatter", "bh_qvalues", "scatter"):
        wrap(ctx.ops, nme, "  bh." + nme)
    wrap(ctx, "_all_gather", "  bh.all_gather")
marks.clear()
barrier()
t0 = time.perf_counter()
for _ in range(reps):
    one_pass()
barrier()
tot = (time.perf_counter() - t0) / reps * 1e3
if rank == 0:
    print("world %d, %d pairs, %d host threads: %.3f ms per pass back to back; run_pass on an idle GPU returns after %.3f ms "
          "(median); %.3f ms per pass with a device sync around every stage" % (world, pairs, E.host_threads(), free,
                                                                               float(np.median(ts)), tot))
    print("  stage timers of the last pass: " + ", ".join("%s %.3f" % kv for kv in stage.items()))
    for k, v in marks.items():
        print("  %-48s %7.3f ms" % (k, v / reps * 1e3))
    print("  %-48s %7.3f ms" % ("unaccounted (python, allocations, launches)",
                                tot - sum(v for k, v in marks.items() if not k.startswith("  ")) / reps * 1e3))
if world > 1:
    dist.destroy_process_group()
