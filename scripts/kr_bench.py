"""Knight-Ruiz balancing of the bench-sized contact matrix on one GPU: time of one product with M + M^T (fhc_kr_spmv) and
of a whole bias computation (fithic_b200/hickry.py).  Usage: python scripts/kr_bench.py [pairs] [file|random]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fithic_b200 import hickry as H, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000_000
order = sys.argv[2] if len(sys.argv) > 2 else "file"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
res = 5000
(m1, m2, c, ch), frags, biases, per = synth.make_intra_device(pairs, res, 1004, dev, mean_count=3.0, with_bias=True, order=order)
off = torch.from_numpy(biases.chr_off).to(dev)
cid = (ch & 0xffff).long()
rows = (off[cid] + m1.long() // res).int()
cols = (off[cid] + m2.long() // res).int()
n = int(biases.chr_off[-1])
kr = H.KRDevice.from_device(rows, cols, c.double(), n)
x = kr.vec(1.0)
y = kr.vec()
for _ in range(3):
    kr.spmv(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 10
for _ in range(reps):
    kr.spmv(x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("fhc_kr_spmv: %d lines, %d loci, line order %s: %.3f ms per product = %.0f GB/s of the 16 B per line" %
      (pairs, n, order, ms, 16 * pairs / ms / 1e6))
rs = y.cpu().numpy()
assert abs(rs.sum() - 2.0 * float(c.double().sum().item())) < 1e-6 * rs.sum()  # row sums of M + M^T = twice the counts
t = time.perf_counter()
removed = H.removeZeroDiagonalCSR(rs, 0.05)
keep = np.ones(n, dtype=bool)
keep[removed] = False
kr.set_kept(np.nonzero(keep)[0])
xk, outer, inner = H.knightRuizAlg(kr)
torch.cuda.synchronize()
dt = time.perf_counter() - t
b = H.computeBiasVector(xk.cpu().numpy()).reshape(-1)
print("whole KR run: %d of %d loci kept, %d outer iterations (last inner %d), %.3f s; bias mean %.4f median %.4f" %
      (kr.n, n, outer, inner, dt, b.mean(), np.median(b)))
