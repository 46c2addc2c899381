"""Multi-GPU parity check (run under torchrun, one rank per GPU): contacts sharded by chromosome, histogram all-reduce,
range-partitioned global BH; every rank compares its shard against the oracle run on the WHOLE data set.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fithic_b200 import synth  # noqa: E402
from fithic_b200.engine import Contacts, Engine, Settings  # noqa: E402
from fithic_b200.parallel import DistCtx  # noqa: E402
from oracle import fithic_oracle as O  # noqa: E402
from tests.util import oracle_inputs, rel_err  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    ctx = DistCtx(device)
    cases = [("intra_bias_p2", dict(n_pairs=400_000, res=100000, seed=77, mean_count=4.0, with_bias=True),
              dict(noOfBins=100, noOfPasses=2), None),
             ("all", dict(n_pairs=300_000, res=100000, seed=78, mean_count=3.0, with_bias=False, inter_fraction=0.3),
              dict(noOfBins=100, allReg=True), None),
             # three passes: the reference's outlier skipping stalls at the first duplicate in FILE order
             ("intra_p3", dict(n_pairs=300_000, res=100000, seed=79, mean_count=4.0, with_bias=True),
              dict(noOfBins=100, noOfPasses=3), None),
             # the range-partitioned BH route forced (FHC_BH_SMALL_SET=0 semantics), and interOnly (BASELINE config 5)
             ("intra_bias_p2_partitioned", dict(n_pairs=400_000, res=100000, seed=77, mean_count=4.0, with_bias=True),
              dict(noOfBins=100, noOfPasses=2), 0),
             ("inter_only", dict(n_pairs=300_000, res=100000, seed=80, mean_count=3.0, with_bias=False, inter_fraction=0.9),
              dict(noOfBins=100, interOnly=True), None),
             ("inter_only_partitioned", dict(n_pairs=300_000, res=100000, seed=80, mean_count=3.0, with_bias=False,
                                             inter_fraction=0.9), dict(noOfBins=100, interOnly=True), 0)]
    for case, kw, sk, small_set in cases:
        contacts, frags, biases, _ = synth.make_intra(**kw)
        st = Settings(resolution=kw["res"], **sk)
        ctx.SMALL_SET = DistCtx.SMALL_SET if small_set is None else small_set
        # shard: intra lines by chromosome (LPT over chromosome sizes), inter lines round-robin by line index
        shards = synth.lpt_shards([int(s) for s in synth.genome(None)[1]], world)
        owner = np.zeros(len(frags.chroms), dtype=np.int64)
        for r, s in enumerate(shards):
            owner[s] = r
        c1 = (contacts.chrs & 0xffff).astype(np.int64)
        c2 = (contacts.chrs >> 16).astype(np.int64)
        line_owner = np.where(c1 == c2, owner[c1], np.arange(len(c1)) % world)
        mine = np.nonzero(line_owner == rank)[0]
        local_c = Contacts(contacts.mid1[mine], contacts.mid2[mine], contacts.cnt[mine], contacts.chrs[mine], contacts.chroms)
        eng = Engine(st, frags, biases, device=device, dist_ctx=ctx)
        eng.upload_contacts(local_c)
        cut = np.flatnonzero(np.diff(mine) != 1) + 1  # this rank's lines as runs of consecutive file lines
        starts = np.concatenate([[0], cut]).astype(np.int64)
        eng.set_line_runs(mine[starts] if len(mine) else np.zeros(0, np.int64),
                          np.diff(np.concatenate([starts, [len(mine)]])) if len(mine) else np.zeros(0, np.int64))
        outl, stats = eng.new_outlier_state()
        oc, fchr, fmid, fh, ost, ob = oracle_inputs(contacts, frags, st, biases)
        want = O.run_pipeline(oc, fchr, fmid, fh, ost, ob)
        for passNo in range(1, st.noOfPasses + 1):
            if passNo > 1 and st.interOnly:
                break
            r = eng.run_pass(passNo, outl, stats)
            torch.cuda.synchronize()
            o = want[passNo - 1]
            assert r["N"] == o["N"] and r["T"] == o["T"], (r["N"], o["N"], r["T"], o["T"])
            assert np.array_equal(r["dists"], o["dists"]) and np.array_equal(r["sums"], o["sums"])
            for i, b in enumerate(o["bins"]):
                assert (int(r["bins"]["lb"][i]), int(r["bins"]["ub"][i]), int(r["bins"]["pairs"][i])) == \
                    (b["lb"], b["ub"], b["pairs"]), (i, b)
            ep = rel_err(r["p"].cpu().numpy(), o["p"][mine])
            eq = rel_err(r["q"].cpu().numpy(), o["q"][mine])
            assert ep <= 1e-6 and eq <= 1e-6, (ep, eq)
            lines = np.repeat(mine, outl.cpu().numpy())
            wl = np.asarray(o["outliersline"], dtype=np.int64)
            assert np.array_equal(lines, wl[np.isin(wl, mine)])
            print("rank %d/%d %s pass %d: lines %d N %d T %d p err %.2e q err %.2e route %s below cut %d" %
                  (rank, world, case, passNo, len(mine), r["N"], r["T"], ep, eq,
                   "gathered" if ctx.last_plan["small_set"] else "partitioned", ctx.last_plan["n_below"]), flush=True)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
