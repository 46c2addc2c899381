"""Where the HOST spends its time in a spline pass: cProfile over a few hundred passes on a small input (the kernels take
microseconds there, what is left is Python, ctypes, torch bookkeeping and the C host stage).  One GPU, or under torchrun."""
import cProfile
import io
import os
import pstats
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
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
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


for _ in range(20):
    one_pass()
torch.cuda.synchronize()
reps = 300
t0 = time.perf_counter()
for _ in range(reps):
    one_pass()
torch.cuda.synchronize()
plain = (time.perf_counter() - t0) / reps * 1e3
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    one_pass()
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    print("world %d, %d pairs: %.3f ms per pass (host stage timers: %s)" % (
        world, pairs, plain, ", ".join("%s %.3f" % (k, v * 1e3) for k, v in eng.timings.get(1, {}).items())))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print("\n".join(ln for ln in s.getvalue().splitlines() if ln.strip())[:6000])
if world > 1:
    ctx.close()
    dist.destroy_process_group()
