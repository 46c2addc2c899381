"""Host-side timeline of one spline pass (wall-clock between synchronisation points), for tuning fixed costs.
Run alone or under torchrun; prints per-stage milliseconds averaged over a few passes on rank 0."""
import os
import sys
import time

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
eng.set_contacts_device(m1, m2, c, ch)

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
wrap(eng, "spline_table", "K2 (eval + host pooling + lut)")
wrap(eng, "pvalues", "K3 (+ lbeta table)")
wrap(eng, "bh_qvalues", "K4 local")
wrap(E, "make_bins", "host make_bins")
wrap(E, "frag_pairs", "host frag_pairs")
wrap(E, "calculate_probabilities", "host probabilities")
wrap(E, "fit_spline", "host spline fit")
if ctx is not None:
    wrap(ctx, "allreduce_hist", "all-reduce hist")
    wrap(ctx, "global_bh", "global BH (exchange + K4)")
    for nme in ("sample_keys", "sort_keys", "partition_count", "partition_scatter", "bh_prepare", "bh_finish", "scatter"):
        wrap(ctx.ops, nme, "  bh." + nme)
    wrap(ctx, "_all_gather", "  bh.all_gather (x3-4)")
reps = 6
for i in range(reps + 2):
    if i == 2:
        marks.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    o, s = eng.new_outlier_state()
    eng.run_pass(1, o, s)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / reps * 1e3
if rank == 0:
    print("world %d: %.2f ms per pass (with a device sync around every stage)" % (world, tot))
    for k, v in marks.items():
        print("  %-38s %7.3f ms" % (k, v / reps * 1e3))
    print("  %-38s %7.3f ms" % ("unaccounted", tot - sum(v for k, v in marks.items() if not k.startswith("  ")) / reps * 1e3))
if world > 1:
    dist.destroy_process_group()
