"""Merge-filter step on one GPU at whole-genome size: time of fhc_merge_components + fhc_merge_select on a synthetic set of
significant bin pairs (clusters around random anchors plus a band along the diagonal, 24 chromosomes), per kernel through the
library's CUDA events.  Usage: python scripts/merge_bench.py [pairs] [--host]   (--host: the serial drivers, for comparison)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fithic_b200 import _capi, merge as G  # noqa: E402


def synthetic_pairs(n, seed=5):
    rng = np.random.default_rng(seed)
    chr_rank = rng.integers(0, 24, n).astype(np.int32)
    nbins = 40_000  # 200 Mb at 5 kb
    na = max(8, n // 20_000)  # anchors per chromosome: ~800 lines around each
    anchors = rng.integers(1, nbins - 400, (24, na))
    a = anchors[chr_rank, rng.integers(0, na, n)] + rng.integers(0, 12, n)
    band = rng.random(n) < 0.3  # close to the diagonal: long thin components
    b = np.where(band, a + rng.integers(1, 4, n), a + rng.integers(1, 12, n) * 16 + rng.integers(0, 8, n))
    cc = rng.integers(2, 200, n).astype(np.int64)
    q = 10.0 ** -rng.uniform(2, 40, n)
    return chr_rank, a.astype(np.int64), b.astype(np.int64), cc, q


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5_000_000
    chr_rank, b1, b2, cc, q = synthetic_pairs(n)
    if "--host" in sys.argv:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from tests.util import merge_components_host
        t = time.perf_counter()
        g = merge_components_host(chr_rank, b1, b2, cc, q, 8, 100, 2, 0)
        dt = time.perf_counter() - t
    else:
        import torch
        G.components_device(chr_rank[:1000], b1[:1000], b2[:1000], cc[:1000], q[:1000], 8, 100, 2, 0)  # warm-up
        _capi.profile_enable(True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        g = G.components_device(chr_rank, b1, b2, cc, q, 8, 100, 2, 0)
        dt = time.perf_counter() - t
        prof = _capi.profile_collect()
        _capi.profile_enable(False)
        for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]:
            print("    %-28s %8.3f ms in %d launch(es)" % (name, v["ms"], v["launches"]))
    roots = np.nonzero(g["label"] == np.arange(n))[0]
    print("%d lines, %d nodes, %d components (largest %d), %d loops kept: %.3f s including the copies to and from the device"
          % (n, int((g["label"] >= 0).sum()), len(roots), int(g["size"][roots].max()), int(g["keep"].sum()), dt))


if __name__ == "__main__":
    main()
