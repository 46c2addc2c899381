"""Debug helper (GPU): list the entries where fhc_bdtrc and the oracle disagree on the p == 1.0 class or by > tol."""
import sys

import numpy as np

sys.path.insert(0, ".")
from fithic_b200 import _capi  # noqa: E402
from oracle import fithic_oracle as O  # noqa: E402
from tests.test_gpu_kernels import gpu_bdtrc  # noqa: E402

lib = _capi.load()
for N in [int(a) for a in sys.argv[1:]] or [5000]:
    rng = np.random.default_rng(N % 9973)
    n = 200_000
    cmax = min(N, 4000)
    cnt = np.minimum(np.floor(np.exp(rng.uniform(0, np.log(cmax + 1), n))).astype(np.int64), cmax)
    ratio = np.exp(rng.uniform(np.log(0.01), np.log(100), n))
    prior = np.minimum(cnt * ratio / N, 1.0)
    prior[::97] = np.exp(rng.uniform(np.log(1e-14), 0, len(prior[::97])))
    want = O.bdtrc(cnt - 1.0, N, prior)
    got = gpu_bdtrc(lib, cnt - 1, N, prior)
    bad = np.nonzero((got == 1.0) != (want == 1.0))[0]
    print("N", N, "class mismatches", len(bad))
    for i in bad[:20]:
        print("  cnt", cnt[i], "prior", repr(prior[i]), "got", repr(got[i]), "want", repr(want[i]))
    with np.errstate(all="ignore"):
        rel = np.abs(got - want) / np.abs(want)
    rel[~np.isfinite(rel)] = 0
    for i in np.argsort(-rel)[:8]:
        print("  rel", rel[i], "cnt", cnt[i], "prior", repr(prior[i]), "got", repr(got[i]), "want", repr(want[i]))
