"""CPU, world_size 2 over gloo: the multi-GPU choreography of fithic_b200/parallel.py (splitters, exchange plan,
all-to-all sizes, carry between key ranges, histogram all-reduce packing) with a numpy stand-in for the CUDA ops.
The stand-in exists only in this test; the product's ops are the kernels of libfithic_b200.so (CudaOps)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fithic_b200 import _capi, parallel
from oracle import fithic_oracle as O

U64_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def key_of(p):
    b = np.asarray(p, dtype=np.float64).view(np.uint64)
    return np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))


class NumpyOps:
    """Same interface as parallel.CudaOps, numpy on CPU tensors (test double)."""

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype)

    def p_cut(self, T, rank_bound):
        return float(_capi.load().fhc_bh_p_cut(float(T), float(rank_bound)))

    def cut_hist(self, p, p_cut0):
        p = p.numpy()
        with np.errstate(invalid="ignore"):
            ok = ~((p == 1.0) | np.isnan(p) | (p >= p_cut0))
        v = p[ok]
        b = np.where(v > 0, np.minimum(v.view(np.uint64) >> np.uint64(47), np.uint64(_capi.BH_CUT_BUCKETS - 1)), 0)
        return torch.from_numpy(np.bincount(b.astype(np.int64), minlength=_capi.BH_CUT_BUCKETS).astype(np.int64))

    def cut_from_hists(self, hists, nranks, my_rank, T, p_cut0):
        lib = _capi.load()
        h = hists.numpy().reshape(nranks, _capi.BH_CUT_BUCKETS)
        tot = np.ascontiguousarray(h.sum(axis=0), dtype=np.uint64)
        cut = float(lib.fhc_host_bh_cut_find(_capi.dptr(tot), float(T), 0.0, float(p_cut0)))
        upto = int(lib.fhc_host_bh_cut_bucket(cut)) if cut < p_cut0 else _capi.BH_CUT_BUCKETS
        share = h[:, :upto].sum(axis=1)
        info = np.zeros(8 + nranks, dtype=np.int64)
        info[:1].view(np.float64)[0] = cut
        info[1], info[2], info[3] = share.sum(), share[my_rank], share.max()
        info[8:] = share
        return torch.from_numpy(info)

    def cut_find(self, hist_host, T, p_cut0):
        h = np.ascontiguousarray(hist_host, dtype=np.uint64)
        return float(_capi.load().fhc_host_bh_cut_find(_capi.dptr(h), float(T), 0.0, float(p_cut0)))

    def sample_keys(self, p, nsamples, p_cut):
        p = p.numpy()
        n = len(p)
        stride = n // nsamples if n > nsamples else 1
        keys = np.full(nsamples, U64_NONE, dtype=np.uint64)
        idx = np.arange(nsamples) * stride
        ok = idx < n
        v = p[idx[ok]]
        k = key_of(v)
        k[(v == 1.0) | np.isnan(v) | (v >= p_cut)] = U64_NONE
        keys[ok] = k
        return torch.from_numpy(keys.view(np.int64))

    def sort_keys(self, keys):
        return torch.from_numpy(np.sort(keys.numpy().view(np.uint64)).view(np.int64))

    def _part(self, p, splitters):
        k = key_of(p)
        return np.searchsorted(np.asarray(splitters, dtype=np.uint64), k, side="right")

    def partition_count(self, p, splitters, p_cut):
        p = p.numpy()
        with np.errstate(invalid="ignore"):
            ok = ~((p == 1.0) | np.isnan(p) | (p >= p_cut))
        return torch.from_numpy(np.bincount(self._part(p[ok], splitters), minlength=len(splitters) + 1).astype(np.int64))

    def partition_scatter(self, p, splitters, send_offsets, q, p_cut, capacity=None, q_prefilled=False):
        p = p.numpy()
        qn = q.numpy()
        with np.errstate(invalid="ignore"):
            one = (p == 1.0) | (p >= p_cut)
        qn[one] = 1.0
        qn[np.isnan(p)] = np.nan
        ok = np.nonzero(~(one | np.isnan(p)))[0]
        part = self._part(p[ok], splitters)
        order = np.argsort(part, kind="stable")
        send = np.zeros(len(p))
        idx = np.zeros(len(p), dtype=np.int32)
        send[:len(ok)] = p[ok][order]
        idx[:len(ok)] = ok[order]
        ends = np.asarray(send_offsets, dtype=np.int64) + np.bincount(part, minlength=len(splitters) + 1)
        self.last_cursors = torch.from_numpy(ends.astype(np.int64))
        return torch.from_numpy(send), torch.from_numpy(idx)

    def bh_qvalues(self, p, T):
        return torch.from_numpy(O.benjamini_hochberg(p.numpy(), T))

    def cut_bucket(self, p_cut):
        return int(_capi.load().fhc_host_bh_cut_bucket(float(p_cut)))

    def bh_prepare(self, p, T, rank_offset, q):
        self._p, self._T = p.numpy().copy(), T
        order = np.argsort(self._p, kind="stable")
        bh = np.minimum(self._p[order] * T / (rank_offset + np.arange(1, len(order) + 1)), 1.0)
        self._order, self._bh = order, bh
        return torch.tensor([bh.max() if len(bh) else 0.0], dtype=torch.float64)

    def bh_finish(self, n, T, rank_offset, floor, q):
        run = np.maximum.accumulate(np.concatenate(([floor], self._bh)))[1:]
        q.numpy()[self._order] = run

    def scatter(self, src, idx, dst):
        dst.numpy()[idx.numpy()] = src.numpy()


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = parallel.DistCtx(device=None, ops=NumpyOps(), samples_per_rank=2048)
        rng = np.random.default_rng(123)
        p_all = rng.random(60_001) ** 3
        p_all[rng.integers(0, len(p_all), 5000)] = 1.0
        p_all[rng.integers(0, len(p_all), 300)] = np.nan
        p_all[rng.integers(0, len(p_all), 4000)] = p_all[rng.integers(0, len(p_all), 4000)]  # ties across ranks
        T = 250_000
        cut = 41_000  # uneven shards
        mine = p_all[:cut] if rank == 0 else p_all[cut:]
        q = torch.full((len(mine),), -1.0, dtype=torch.float64)
        want = O.benjamini_hochberg(p_all, T)
        want = want[:cut] if rank == 0 else want[cut:]
        # few survivors: every rank ranks the gathered set itself
        ctx.global_bh(None, torch.from_numpy(mine.copy()), float(T), q=q)
        assert ctx.last_plan["small_set"]
        assert np.array_equal(q.numpy(), want, equal_nan=True), "global BH (small set) differs on rank %d" % rank
        # many survivors: range partition over the ranks
        ctx.SMALL_SET = 100
        q = torch.full((len(mine),), -1.0, dtype=torch.float64)
        ctx.global_bh(None, torch.from_numpy(mine.copy()), float(T), q=q)
        assert not ctx.last_plan["small_set"]
        assert np.array_equal(q.numpy(), want, equal_nan=True), "global BH differs on rank %d" % rank
        cm = ctx.last_plan["count_matrix"]
        with np.errstate(invalid="ignore"):
            assert cm.sum() == np.sum(~((p_all == 1.0) | np.isnan(p_all) | (p_all >= ctx.last_plan["p_cut"])))
        assert 0.2 < ctx.last_plan["p_cut0"] < 0.3  # 60,001 lines / T = 250,000: the rank bound alone drops three quarters
        assert ctx.last_plan["p_cut"] <= ctx.last_plan["p_cut0"]
        assert abs(cm[:, 0].sum() - cm[:, 1].sum()) < 0.1 * cm.sum()  # the sample balances the two key ranges

        assert ctx.last_plan["n_below"] == cm.sum()
        # nothing below the cut: no exchange at all
        ones = np.where(rng.random(5000) < 0.3, 1.0, 0.9 + 0.1 * rng.random(5000))
        q = torch.full((5000,), -1.0, dtype=torch.float64)
        ctx.global_bh(None, torch.from_numpy(ones.copy()), 20000.0, q=q)
        assert ctx.last_plan["n_below"] == 0 and np.all(q.numpy() == 1.0)

        # exchange 1: [histogram | totals | one slot per rank for the largest count] in one all-reduce; the seen bits are
        # only exchanged when some rank counted lines with a count <= 0
        D = 70
        ns = _capi.N_SCALARS + world
        fused = torch.zeros(D + ns, dtype=torch.int64)
        hist, scal = fused[:D], fused[D:]
        present = torch.zeros((D + 31) // 32, dtype=torch.int32)
        hist[rank::2] = rank + 1
        seen = [3, 40] if rank == 0 else [40, 63, 69]
        words = np.zeros((D + 31) // 32, dtype=np.uint32)
        for b in seen:
            words[b >> 5] |= np.uint32(1 << (b & 31))
        present.copy_(torch.from_numpy(words.view(np.int32)))
        scal[_capi.S_INTRA_INRANGE_SUM] = 100 + rank
        scal[_capi.S_NONPOS_LINES] = len(seen)
        scal[_capi.N_SCALARS + rank] = 7 if rank == 0 else 19  # what fhc_hist_distance does with n_rank_slots = world
        ctx.allreduce_k1(fused)
        ctx.or_present(present)
        exp = np.zeros(D, dtype=np.int64)
        exp[0::2] = 1
        exp[1::2] = 2
        assert np.array_equal(hist.numpy(), exp)
        got_bits = np.unpackbits(present.numpy().view(np.uint8), bitorder="little")[:D]
        assert sorted(np.nonzero(got_bits)[0].tolist()) == [3, 40, 63, 69]
        assert int(scal[_capi.S_INTRA_INRANGE_SUM]) == 201 and int(scal[_capi.S_NONPOS_LINES]) == 5
        assert scal[_capi.N_SCALARS:].tolist() == [7, 19] and int(scal[_capi.S_MAX_COUNT]) == 0
        assert ctx.n_global(10 + rank) == 21 and ctx.n_global(10 + rank) == 21
        # the host stage shares its possible-pair sums through shared memory: same bits as one rank on its own
        import ctypes
        from fithic_b200 import synth
        from tests.test_host_stage import _stage_io, _views
        lib = _capi.load()
        assert ctx.shm is not None
        contacts, frags, _, _ = synth.make_intra(60_000, 100000, seed=5, mean_count=4.0)
        d = np.abs(contacts.mid1.astype(np.int64) - contacts.mid2)
        D = int(max(d.max(), frags.max_mid.max()) // 100000 + 2)
        hist = np.bincount(d // 100000, weights=contacts.cnt, minlength=D).astype(np.int64)
        scal = np.zeros(_capi.N_SCALARS, dtype=np.uint64)
        scal[_capi.S_INTRA_INRANGE_SUM] = int(hist.sum())
        scal[_capi.S_MAX_COUNT] = int(contacts.cnt.max())
        io1, keep1 = _stage_io(lib, hist, scal, None, 100000, 100, frags, 0, -1, 1, 1)
        _capi.check(lib.fhc_host_stage(ctypes.byref(io1), 7))
        io2, keep2 = _stage_io(lib, hist, scal, None, 100000, 100, frags, 0, -1, 1, 2)
        io2.pairs_rank, io2.pairs_world, io2.shm = rank, world, ctx.shm
        for _ in range(3):  # several passes: the slots alternate
            _capi.check(lib.fhc_host_stage(ctypes.byref(io2), 7))
            a, b = _views(io1, keep1, 100), _views(io2, keep2, 100)
            for k in ("sumdist", "x_bins", "y_bins", "t", "c", "table", "lut", "pairs"):
                assert np.array_equal(a[k], b[k]), k
        assert ctx.max_int(5 + rank) == 6
        assert ctx.allreduce_small(np.array([1, 2 + rank])).tolist() == [2, 5]
        open(os.path.join(tmp, "ok%d" % rank), "w").close()
    finally:
        try:
            ctx.close()
        except Exception:
            pass
        dist.destroy_process_group()


def test_two_rank_choreography(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_plan_helpers():
    keys = np.sort(np.concatenate([np.arange(1000, dtype=np.uint64), np.full(24, U64_NONE)]))
    sp = parallel.choose_splitters(keys, 4)
    assert sp.tolist() == [250, 500, 750]
    assert parallel.choose_splitters(np.full(8, U64_NONE), 4).tolist() == [0, 0, 0]
    cm = np.array([[5, 1, 0], [2, 2, 2], [0, 7, 1]])
    send, recv, off, soff = parallel.exchange_plan(cm, 1)
    assert (send, recv, off, soff.tolist()) == ([2, 2, 2], [1, 2, 7], 7, [0, 2, 4])
    assert parallel.exchange_plan(cm, 0)[2] == 0 and parallel.exchange_plan(cm, 2)[2] == 17
    assert parallel.carry_floor([0.3, 0.9, 0.5], 0) == 0.0
    assert parallel.carry_floor([0.3, 0.9, 0.5], 2) == 0.9
    lib = _capi.load()
    for v in (0.0, 1e-300, 0.25, 0.5, 0.999999, 1.0, 2.5):
        assert lib.fhc_bh_key_of(v) == int(key_of([v])[0])
