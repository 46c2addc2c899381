"""Multi-GPU execution: one process per GPU, contacts sharded by chromosome, torch.distributed (NCCL) as the bootstrap and fall-back, the library's own
collectives over CUDA IPC peer windows (csrc/comm.cu, NVLink / NVSwitch) for the two exchange steps of a spline pass.

  exchange 1  ONE all-reduce (sum) of [distance histogram | observed totals | one slot per rank for the largest count]
              (<= 400 kB, latency bound).  Every rank then runs the identical host binning / spline fit on identical
              inputs, so tables are bit-identical everywhere.
  exchange 2  global BH: ONE all-reduce of the ranks' 32768-bucket value histograms; a kernel finds the cut above which
              every q is 1.0 and the number of p-values below it (on all ranks, on this one), and ONE small read-back tells
              the host how to go on: nothing below the cut (sparse maps without signal) -> done; few -> all-gather of the
              survivors, every rank ranks the small global set itself; many -> the p-values below the cut are
              range-partitioned by value so that rank r ranks one contiguous key range:
              sample keys -> all-gather -> splitters; count per part -> all-gather -> offsets; scatter into send
              buffers (kernel) -> all-to-all(v) of p -> local compaction/sort/tile maxima (kernels) -> all-gather of the
              per-range maxima (carry) -> scan + scatter (kernels) -> all-to-all(v) of q back -> scatter to line order.

The collective choreography only needs a small "ops" interface (sample / count / scatter / prepare / finish / ...),
implemented by the CUDA library in production (CudaOps) -- world_size-2 gloo tests drive the same choreography with a
numpy stand-in to check the bookkeeping (offsets, splits, carries) without a GPU.
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import check, dptr

_U64_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def choose_splitters(sorted_sample_keys, nparts):
    """Splitter keys at the k/nparts quantiles of the valid (non-UINT64_MAX) sample keys; ascending, length nparts-1."""
    s = np.asarray(sorted_sample_keys, dtype=np.uint64)
    s = s[s != _U64_NONE]
    if len(s) == 0:
        return np.zeros(max(nparts - 1, 0), dtype=np.uint64)
    pos = (np.arange(1, nparts) * len(s)) // nparts
    return s[np.minimum(pos, len(s) - 1)].astype(np.uint64)


def exchange_plan(count_matrix, rank):
    """count_matrix[src, part] = rankable p-values src sends to part.  Returns (send_splits, recv_splits, rank_offset,
    send_offsets) for `rank`: rank_offset = number of keys in all lower parts (global rank of this part's first key)."""
    cm = np.asarray(count_matrix, dtype=np.int64)
    send = cm[rank].tolist()
    recv = cm[:, rank].tolist()
    rank_offset = int(cm[:, :rank].sum())
    send_off = np.concatenate([[0], np.cumsum(cm[rank])[:-1]]).astype(np.int64)
    return send, recv, rank_offset, send_off


def carry_floor(local_maxima, rank):
    """Running max handed to `rank`: the largest bh value of every lower key range (0 for the first, myStats.py:30)."""
    m = 0.0
    for v in list(local_maxima)[:rank]:
        m = max(m, float(v))
    return m


class CudaOps:
    """The kernels of libfithic_b200.so behind the ops interface (device tensors in, device tensors out)."""

    def __init__(self, device):
        self.lib = _capi.load()
        self.device = device
        self._ws = None
        self._cursor1 = None

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]

    def p_cut(self, T, rank_bound):
        return float(self.lib.fhc_bh_p_cut(float(T), float(rank_bound)))

    def cut_hist(self, p, p_cut0):
        """Value histogram of the rankable p-values below p_cut0 (int64 [BH_CUT_BUCKETS], device)."""
        hist = torch.zeros(_capi.BH_CUT_BUCKETS, dtype=torch.int64, device=self.device)
        check(self.lib.fhc_bh_cut_hist(dptr(p), p.numel(), float(p_cut0), dptr(hist), self._stream()))
        return hist

    def cut_from_hists(self, hists, nranks, my_rank, T, p_cut0):
        """Every rank's histogram (nranks x BH_CUT_BUCKETS, device) -> info block on the device (see bh.cu)."""
        info = self.empty(8 + nranks, torch.int64)
        check(self.lib.fhc_bh_cut_from_hists(dptr(hists), int(nranks), int(my_rank), float(T), float(p_cut0), dptr(info),
                                             self._stream()))
        return info

    def cut_find(self, hist_host, T, p_cut0):
        h = np.ascontiguousarray(hist_host, dtype=np.uint64)
        return float(self.lib.fhc_host_bh_cut_find(dptr(h), float(T), 0.0, float(p_cut0)))

    def sample_keys(self, p, nsamples, p_cut):
        keys = self.empty(nsamples, torch.int64)
        check(self.lib.fhc_bh_sample_keys(dptr(p), p.numel(), nsamples, float(p_cut), dptr(keys), self._stream()))
        return keys

    def sort_keys(self, keys):
        n = keys.numel()
        vals = self.empty(n, torch.int32)
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        wsb = int(self.lib.fhc_sort_workspace_bytes(n))
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        check(self.lib.fhc_sort_pairs_u64(dptr(keys), dptr(vals), dptr(ko), dptr(vo), n, dptr(ws), wsb, self._stream()))
        return ko

    def partition_count(self, p, splitters, p_cut):
        nparts = len(splitters) + 1
        counts = self.empty(nparts, torch.int64)
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        check(self.lib.fhc_bh_partition_count(dptr(p), p.numel(), dptr(sp), nparts, float(p_cut), dptr(counts),
                                              self._stream()))
        return counts

    def partition_scatter(self, p, splitters, send_offsets, q, p_cut, capacity=None, q_prefilled=False):
        nparts = len(splitters) + 1
        n = p.numel()
        if nparts == 1 and int(send_offsets[0]) == 0:  # one part from slot 0: no host->device copy
            if self._cursor1 is None:
                self._cursor1 = torch.zeros(1, dtype=torch.int64, device=self.device)
            else:
                self._cursor1.zero_()
            cursors = self._cursor1
        else:
            cursors = torch.from_numpy(np.ascontiguousarray(send_offsets, dtype=np.int64)).to(self.device)
        cap = n if capacity is None else int(capacity)  # the caller knows how many p-values lie below the cut
        send = torch.empty(max(cap, 1), dtype=torch.float64, device=self.device)  # (never empty: a null pointer is refused)
        idx = torch.empty(max(cap, 1), dtype=torch.int32, device=self.device)
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        check(self.lib.fhc_bh_partition_scatter(dptr(p), n, dptr(sp), nparts, float(p_cut), dptr(cursors), dptr(send),
                                                dptr(idx), dptr(q), 1 if q_prefilled else 0, self._stream()))
        self.last_cursors = cursors  # after the call: one past the last slot used by each part
        return send, idx

    def _bh_ws(self, n):
        wsb = int(self.lib.fhc_bh_workspace_bytes(n))
        if self._ws is None or self._ws.numel() < wsb:
            self._ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        return self._ws, wsb

    def bh_prepare(self, p, T, rank_offset, q):
        n = p.numel()
        ws, wsb = self._bh_ws(n)
        local_max = self.empty(1, torch.float64)
        # the received p-values are already below the global p_cut: rank all of them
        check(self.lib.fhc_bh_prepare(dptr(p), n, float(T), int(rank_offset), float("inf"), dptr(q), dptr(local_max), None,
                                      dptr(ws), wsb, self._stream()))
        return local_max

    def bh_finish(self, n, T, rank_offset, floor, q):
        ws, wsb = self._bh_ws(n)
        check(self.lib.fhc_bh_finish(n, float(T), int(rank_offset), float(floor), dptr(q), dptr(ws), wsb, self._stream()))

    def scatter(self, src, idx, dst):
        check(self.lib.fhc_scatter_f64(dptr(src), dptr(idx), src.numel(), dptr(dst), self._stream()))

    def bh_qvalues(self, p, T):
        """q-values of one array on this GPU (the whole of K4)."""
        n = p.numel()
        q = self.empty(n, torch.float64)
        ws, wsb = self._bh_ws(n)
        check(self.lib.fhc_bh_qvalues(dptr(p), n, float(T), 0, 0.0, dptr(q), None, None, dptr(ws), wsb, self._stream()))
        return q

    def cut_bucket(self, p_cut):
        return int(self.lib.fhc_host_bh_cut_bucket(float(p_cut)))


class DistCtx:
    """Collective choreography of one rank.  `ops` defaults to the CUDA library."""

    SMALL_SET = 1 << 22  # at most this many p-values below the cut: every GPU ranks the gathered set itself

    def __init__(self, device=None, ops=None, group=None, samples_per_rank=1 << 16):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device
        self.ops = ops if ops is not None else CudaOps(device)
        self.samples_per_rank = samples_per_rank
        self._n_global = None
        # The two small exchanges of a pass go through the library's own NVLink collectives (csrc/comm.cu: peer windows
        # shared by CUDA IPC, two kernels per collective) when every rank can set them up; FHC_COMM=nccl keeps NCCL.
        self.comm = None
        self.comm_slot_bytes = 0
        self._cut_work = self._cut_info = None
        if isinstance(self.ops, CudaOps) and self.world > 1 and os.environ.get("FHC_COMM", "p2p") != "nccl":
            self._init_p2p()
        # host-side exchange (POSIX shared memory) for the sums the host stage shares among the ranks of a node
        self.shm = None
        if self.world > 1 and os.environ.get("FHC_PAIRS_SPLIT", "1") != "0":
            self._init_shm()
        env = os.environ.get("FHC_BH_SMALL_SET")  # e.g. 0: always take the range-partitioned route (tests, timing)
        if env is not None:
            self.SMALL_SET = int(env)
        self.pinned_cores = None
        if isinstance(self.ops, CudaOps) and os.environ.get("FHC_PIN_CORES", "0") == "1":
            self._pin_cores()

    def _pin_cores(self):
        """Give this rank its own slice of the node's cores (FHC_PIN_CORES=1; off by default).  Measured at 8 GPUs on a
        32-core box: the waits for the slowest rank at the two exchanges of a pass go away (0.14 -> 0.04 ms), but every
        rank's host stage gets slower inside its four cores (fit 0.17 -> 0.24 ms), and the pass takes the same 1.9 ms."""
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
        local_rank = int(os.environ.get("LOCAL_RANK", "0") or 0)
        try:
            cores = sorted(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            return
        per = len(cores) // max(local_world, 1)
        if local_world <= 1 or per < 2:
            return
        mine = cores[local_rank * per:(local_rank + 1) * per]
        try:
            os.sched_setaffinity(0, mine)  # the calling thread; the host pool's workers are created later and inherit it
            self.pinned_cores = mine
            os.environ["FHC_CORES_PINNED"] = "1"  # host_threads(): the affinity mask is this rank's own slice now
        except OSError:
            pass

    def _init_p2p(self, slot_bytes=1 << 20):
        lib = self.ops.lib
        ok, comm = 1, ctypes.c_void_p()
        hb = int(lib.fhc_comm_handle_bytes())
        mine = ctypes.create_string_buffer(hb)
        if self.world > 16 or lib.fhc_comm_create(self.rank, self.world, slot_bytes, ctypes.byref(comm), mine) != 0:
            ok = 0
        t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(self.device)
        allh = self._all_gather(t).cpu().numpy().tobytes()
        if ok and lib.fhc_comm_connect(comm, allh) != 0:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int64, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)  # all ranks or none
        if int(flag.item()) == 1:
            self.comm, self.comm_slot_bytes = comm, slot_bytes
        elif comm:
            lib.fhc_comm_destroy(comm)
        dist.barrier(group=self.group)

    def _init_shm(self, slot_bytes=1 << 16):
        """Rank 0 names and creates the shared-memory object, the others open it (one node: torchrun --nnodes=1)."""
        import time
        lib = _capi.load()
        dev = self.device if self.device is not None else "cpu"
        name = torch.zeros(64, dtype=torch.uint8, device=dev)
        if self.rank == 0:
            raw = ("/fhc_b200_%d_%d" % (os.getpid(), time.time_ns() % 10 ** 12)).encode()
            name[:len(raw)] = torch.tensor(list(raw), dtype=torch.uint8)
        dist.broadcast(name, src=0, group=self.group)
        raw = bytes(name.cpu().tolist()).rstrip(b"\0")
        handle = ctypes.c_void_p()
        ok = 1 if lib.fhc_shm_open(raw, self.rank, self.world, slot_bytes, ctypes.byref(handle)) == 0 else 0
        flag = torch.tensor([ok], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)  # all ranks or none
        if int(flag.item()) == 1:
            self.shm = handle
        elif ok:
            lib.fhc_shm_close(handle)

    def close(self):
        if self.comm is not None:
            self.ops.lib.fhc_comm_destroy(self.comm)
            self.comm = None
        if getattr(self, "shm", None) is not None:
            _capi.load().fhc_shm_close(self.shm)
            self.shm = None

    # ---- exchange 1 -------------------------------------------------------------------------------------------------
    def allreduce_k1(self, fused):
        """Sum K1's [hist | totals | rank slots] over the ranks, in place: one collective, no host synchronisation (the
        largest count travels in the rank slots, see fhc_hist_distance)."""
        n = fused.numel()
        if self.comm is not None and n % 2 == 0 and 8 * n <= self.comm_slot_bytes and fused.dtype == torch.int64:
            check(self.ops.lib.fhc_comm_allreduce_u64(self.comm, dptr(fused), n, self.ops._stream()))
        else:
            dist.all_reduce(fused, op=dist.ReduceOp.SUM, group=self.group)

    def or_present(self, present):
        """OR the 'distance seen with counts <= 0' bitmaps of all ranks, in place.  Only needed when the summed totals
        say that such lines exist at all (scalars[S_NONPOS_LINES] != 0): rare."""
        g = self._all_gather(present).view(self.world, -1)
        ored = g[0]
        for w in range(1, self.world):
            ored = torch.bitwise_or(ored, g[w])
        present.copy_(ored)

    def max_int(self, v):
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def min_int(self, v):
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return int(t.item())

    def allreduce_small(self, arr):
        """Sum a small host int64 array over ranks (per-bin outlier decrements)."""
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    # ---- exchange 2 -------------------------------------------------------------------------------------------------
    def _all_gather(self, t):
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        nbytes = t.numel() * t.element_size()
        if self.comm is not None and nbytes % 16 == 0 and 0 < nbytes <= self.comm_slot_bytes and t.is_contiguous() \
                and t.data_ptr() % 16 == 0:
            check(self.ops.lib.fhc_comm_allgather(self.comm, dptr(t), dptr(out), nbytes, self.ops._stream()))
        else:
            dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def n_global(self, n):
        """Lines of the whole file (sum over ranks); the shard sizes do not change between passes, so one collective per
        shard size."""
        if self._n_global is None or self._n_global[0] != n:
            t = torch.tensor([n], dtype=torch.int64, device=self.device if self.device is not None else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._n_global = (n, int(t.item()))
        return self._n_global[1]

    def global_bh(self, engine, p, T, q=None, q_prefilled=False):
        """q-values of the union of every rank's p-values (myStats.benjamini_hochberg_correction over the whole file).
        q_prefilled: q holds 1.0 everywhere already, so only the other values are written."""
        ops, G, r = self.ops, self.world, self.rank
        n = p.numel()
        if q is None:
            q = engine._tensor("q", n, torch.float64) if engine is not None else ops.empty(n, torch.float64)
        # 0. p-values that are certain to end with q = 1.0 are neither exchanged nor ranked (bh.cu: bh_p_cut) ...
        p_cut0 = ops.p_cut(T, self.n_global(n))
        #    ... and the value histograms of what is left say where q reaches 1.0 for good (bh.cu: cut_bucket_closes): one
        #    all-gather of the histograms, one kernel, one read-back of a few numbers
        counts = None
        if self.comm is not None:  # histogram, all-reduce, cut kernel and read-back in one library call
            if self._cut_work is None:
                self._cut_work = torch.empty(2 * _capi.BH_CUT_BUCKETS + 8, dtype=torch.int64, device=self.device)
                self._cut_info = torch.empty(8, dtype=torch.int64).pin_memory()
            check(ops.lib.fhc_bh_dist_cut(self.comm, dptr(p), n, float(T), float(p_cut0), dptr(self._cut_work),
                                          dptr(self._cut_info), dptr(q) if q_prefilled else None, ops._stream()))
            info = self._cut_info.numpy()
            p_cut = float(info[:1].view(np.float64)[0])
            n_below, mine = int(info[1]), int(info[2])
            mx = mine
            if n_below == 0 and q_prefilled:  # q is final: 1.0 from the fill, NaN from the histogram sweep
                self.last_plan = dict(splitters=np.zeros(0, dtype=np.uint64), count_matrix=np.zeros((G, 1), dtype=np.int64),
                                      rank_offset=0, floor=0.0, p_cut=p_cut, p_cut0=p_cut0, small_set=True, n_below=0)
                return q
            if 0 < n_below <= self.SMALL_SET:  # every rank's share, for the sizes of the exchange
                counts = self._all_gather(torch.tensor([mine, 0], dtype=torch.int64, device=p.device)).cpu().numpy()[0::2]
                mx = int(counts.max())
        else:
            hist = ops.cut_hist(p, p_cut0)
            info = ops.cut_from_hists(self._all_gather(hist), G, r, T, p_cut0).cpu().numpy()
            p_cut = float(info[:1].view(np.float64)[0])
            n_below, mine, mx = int(info[1]), int(info[2]), int(info[3])
            counts = info[8:8 + G].astype(np.int64)
        if counts is None:
            counts = np.zeros(G, dtype=np.int64)
        # 0b. few survivors (the usual case on a sparse map): no range partition -- every rank compacts its survivors,
        #     the survivors are all-gathered (padded to the largest share), every rank ranks the small global set itself
        #     and keeps the q-values of its own lines.  Nothing below the cut: one kernel writes q = 1 / NaN and that is it.
        if n_below <= self.SMALL_SET:
            send, idx = ops.partition_scatter(p, np.zeros(0, dtype=np.uint64), np.zeros(1, dtype=np.int64), q, p_cut,
                                              capacity=mine, q_prefilled=q_prefilled)
            if n_below:
                pad = ops.empty(mx, torch.float64)
                pad[:mine] = send[:mine]
                if mine < mx:
                    pad[mine:] = 1.0
                allp = self._all_gather(pad).view(G, mx)
                glob = torch.cat([allp[g, :int(counts[g])] for g in range(G)]) if G > 1 else allp[0, :mine]
                qg = ops.bh_qvalues(glob.contiguous(), T)
                off = int(counts[:r].sum())
                if mine:
                    ops.scatter(qg[off:off + mine].contiguous(), idx[:mine], q)
            self.last_plan = dict(splitters=np.zeros(0, dtype=np.uint64), count_matrix=counts.reshape(G, 1), rank_offset=0,
                                  floor=0.0, p_cut=p_cut, p_cut0=p_cut0, small_set=True, n_below=n_below)
            return q
        # 1. splitters from a sorted sample of everybody's keys
        sample = ops.sample_keys(p, self.samples_per_rank, p_cut)
        allsamp = ops.sort_keys(self._all_gather(sample))
        splitters = choose_splitters(allsamp.cpu().numpy().view(np.uint64), G)
        # 2. how many keys go where
        counts = ops.partition_count(p, splitters, p_cut)
        cm = self._all_gather(counts).cpu().numpy().reshape(G, G)
        send_splits, recv_splits, rank_offset, send_off = exchange_plan(cm, r)
        # 3. group by destination, exchange
        send, idx = ops.partition_scatter(p, splitters, send_off, q, p_cut, capacity=mine, q_prefilled=q_prefilled)
        n_send, n_recv = int(sum(send_splits)), int(sum(recv_splits))
        recv = ops.empty(n_recv, torch.float64)
        dist.all_to_all_single(recv, send[:n_send], recv_splits, send_splits, group=self.group)
        # 4. rank my key range; the carry from lower ranges arrives between the two halves
        q_recv = ops.empty(n_recv, torch.float64)
        local_max = ops.bh_prepare(recv, T, rank_offset, q_recv)
        floor = carry_floor(self._all_gather(local_max).cpu().numpy(), r)
        ops.bh_finish(n_recv, T, rank_offset, floor, q_recv)
        # 5. send the q-values home and put them back in line order
        q_back = ops.empty(n_send, torch.float64)
        dist.all_to_all_single(q_back, q_recv, send_splits, recv_splits, group=self.group)
        ops.scatter(q_back, idx[:n_send], q)
        self.last_plan = dict(splitters=splitters, count_matrix=cm, rank_offset=rank_offset, floor=floor, p_cut=p_cut,
                              p_cut0=p_cut0, small_set=False, n_below=n_below)
        return q
