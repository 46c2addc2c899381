"""Time the spline pass on the bench input under the library's experiment switches (read per call from the environment):
per-kernel device ms for each variant.  One process, one data set.  Usage: python scripts/k3_variants.py [pairs]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fithic_b200 import _capi, synth  # noqa: E402
from fithic_b200.engine import Engine, Settings  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000_000
order = sys.argv[2] if len(sys.argv) > 2 else "file"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
(m1, m2, c, ch), frags, biases, per = synth.make_intra_device(pairs, 5000, 1004, dev, mean_count=3.0, with_bias=True, order=order)
st = Settings(resolution=5000, noOfBins=100)
eng = Engine(st, frags, biases, device=dev)
# chromosome ids as runs, as bench.py and the CLI pass them (K1 and K3 then read no chrs array)
mine = [k for k in range(len(per)) if per[k] > 0]
runs = (np.array([k | (k << 16) for k in mine], dtype=np.uint32), np.array([per[k] for k in mine], dtype=np.int64))
eng.set_contacts_device(m1, m2, c, ch, chr_runs=None if os.environ.get("K3V_CHRS") else runs)
VARIANTS = [
    ("lists: front v2 (default)", {}),
    ("lists: front v1", {"FHC_PVAL_FRONT": "v1"}),
]
KEYS = ("FHC_PVAL_FRONT", "FHC_PVAL_FINISH_OCC", "FHC_PVAL_IMPL", "FHC_BH_TIGHTEN", "FHC_PREPASS", "FHC_PVAL_FINISH")
if os.environ.get("K3V_ONLY"):
    VARIANTS = VARIANTS[:int(os.environ["K3V_ONLY"])]
print("line order:", order)
ref = None
for name, env in VARIANTS:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(2):
        o, s = eng.new_outlier_state()
        r = eng.run_pass(1, o, s)
    torch.cuda.synchronize()
    _capi.profile_enable(True)
    _capi.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    e0.record()
    for _ in range(reps):
        o, s = eng.new_outlier_state()
        r = eng.run_pass(1, o, s)
    e1.record()
    torch.cuda.synchronize()
    prof = _capi.profile_collect()
    _capi.profile_enable(False)
    p = r["p"].clone()
    if ref is None:
        ref = p
    diff = float(((p - ref).abs() / ref.abs().clamp_min(1e-300)).nan_to_num(0).max().item())
    k3 = sum(v["ms"] for k, v in prof.items() if k.startswith("pval")) / reps
    print("%-28s step %.2f ms  K3 %.2f ms  max rel diff of p vs first variant %.2e" % (name, e0.elapsed_time(e1) / reps, k3, diff))
    print("    " + ", ".join("%s %.3f" % (k.replace("_kernel", ""), v["ms"] / reps)
                             for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:9]))
