"""SASS evidence for the main kernels (run on the build box): per kernel of libfithic_b200.so the registers, and how many
128-bit / 64-bit global loads and stores (streaming variants .NA / .EF included), shared-memory atomics, global reductions, warp votes / matches / shuffles and FP64
FMAs its code holds.  Usage: python scripts/sass_summary.py [kernel-regex] > profiles/sass_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "fithic_b200", "libfithic_b200.so")
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                  r"hist_distance|pval_front|pval_iterate|pval_finish|pval_prepass|bh_compact|bh_cut_hist|radix_onesweep|"
                  r"bh_scatter|fill_f64|comm_push|comm_sum|frag_pairs|outlier_bin|digest")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)[^\n]*SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
keys = [("LDG.128", r"LDG\.E(\.[A-Z]+)*\.128"), ("LDG.64", r"LDG\.E(\.[A-Z]+)*\.64"),
        ("LDG other", r"LDG\.E(?!(\.[A-Z]+)*\.(128|64))"),
        ("STG.128", r"STG\.E(\.[A-Z]+)*\.128"), ("STG.64", r"STG\.E(\.[A-Z]+)*\.64"), ("ATOMS", r"ATOMS"), ("RED/ATOMG", r"\b(RED|ATOMG|ATOM)\b"),
        ("MATCH", r"MATCH"), ("VOTE", r"VOTE"), ("SHFL", r"SHFL"), ("DFMA", r"DFMA"), ("DMUL", r"DMUL"), ("MUFU", r"MUFU"),
        ("IMAD.HI", r"IMAD\.HI"), ("BAR", r"BAR\.SYNC")]
print("cuobjdump -sass / -res-usage of fithic_b200/libfithic_b200.so (sm_100a): static instruction counts per kernel")
print("%-58s %4s %7s " % ("kernel", "regs", "smem") + " ".join("%9s" % k for k, _ in keys) + "  total")
cur, counts = None, None
out = []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            out.append((cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts["total"] += 1
        for k, pat in keys:
            if re.search(pat, line):
                counts[k] += 1
if cur:
    out.append((cur, counts))
for name, c in out:
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    dem = dem.replace("fhc::", "")
    if not want.search(dem):
        continue
    r = regs.get(name, (0, 0))
    print("%-58s %4d %7d " % (dem[:58], r[0], r[1]) + " ".join("%9d" % c[k] for k, _ in keys) + "  %5d" % c["total"])
