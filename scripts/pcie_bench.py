"""What the host side of this box can move: every rank copies pinned host memory to its GPU and back, alone and all ranks at
once (the end-to-end leg of bench.py does exactly that with the contacts and the results of a pass).  Under torchrun."""
import os
import sys
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(0)
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.ones(n, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=4):
    fn()
    barrier()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    barrier()
    return dt


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    s2.synchronize()


res = {}
for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both directions", both, 2 * n)):
    # all ranks at once
    dt = timed(fn)
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name + " all ranks"] = nbytes * world / t.item() / 1e9
    # rank 0 alone
    if rank == 0:
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        res[name + " rank 0 alone"] = nbytes / ((time.perf_counter() - t0) / 4) / 1e9
    barrier()
if rank == 0:
    print("world %d, %d MiB per copy, pinned host memory; GB/s (aggregate over ranks where all ranks copy at once)" % (world, mb))
    for k, v in res.items():
        print("  %-32s %8.1f" % (k, v))
if world > 1:
    dist.destroy_process_group()
