"""Dynamic SASS opcode mix of one kernel from the source page of an .ncu-rep (instructions executed per opcode, per unit of
work), heaviest first.  Run on the CPU box.
Usage: python scripts/ncu_opcode_mix.py REPORT.ncu-rep KERNEL_REGEX UNITS_PER_LAUNCH [top_n]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = rows[0][1]
heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
launches = len(heads)
hdr = rows[heads[0]]
S, I, T, N = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
warp, thread, samp = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows:
    if len(r) <= max(S, I, T, N) or not r[I].isdigit():
        continue
    text = re.sub(r"^@!?U?P\d+\s+", "", r[S].strip())
    op = text.split()[0].split(".")[0] if text else "?"
    warp[op] += int(r[I])
    thread[op] += int(r[T])
    samp[op] += int(r[N])
tw, tt, ts = sum(warp.values()), sum(thread.values()), max(sum(samp.values()), 1)
print("%s\n%d launch(es) in the report; per launch %.4g warp instructions, %.1f thread instructions per unit of work (%g units)"
      % (name, launches, tw / launches, tt / launches / units, units))
print("%-8s %8s %14s %9s" % ("opcode", "warp %", "thread inst/unit", "samples %"))
for op, n in warp.most_common(top):
    print("%-8s %7.2f%% %14.2f %8.1f%%" % (op, 100.0 * n / tw, thread[op] / launches / units, 100.0 * samp[op] / ts))
