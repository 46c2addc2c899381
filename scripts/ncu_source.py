"""Per-CUDA-source-line cost of one kernel: joins the SASS page of an .ncu-rep (instructions executed, stall samples, active
threads per instruction) with the line table of the object file the report was taken from (nvdisasm -g), heaviest lines
first.  Run on the CPU box; the object must be the build that ran.
Usage: python scripts/ncu_source.py REPORT.ncu-rep OBJECT.o KERNEL_REGEX [top_n] [exec]   (exec: heaviest by instructions
executed instead of by stall samples)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by_exec = len(sys.argv) > 5 and sys.argv[5] == "exec"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = rows[0][1]
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
ix = {k: hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
data = [r for r in rows[hi + 1:] if len(r) > max(ix.values()) and r[ix["Source"]] != "Source"]
# line table
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
short = re.sub(r"\(.*", "", name).replace("void ", "").replace("fhc::", "")
short = re.sub(r"<.*", "", short)
lines, cur, inside = [], "?", False
paths = {}
targs = re.search(r"<([^>]*)>", name.split("(fhc::")[0] if "(fhc::" in name else name)
mangled_args = ""
if targs:  # (bool)1, (int)4 -> ILb1ELi4EE
    for a in targs.group(1).split(","):
        m = re.match(r"\s*\((bool|int)\)(-?\d+)", a)
        if m:
            mangled_args += ("Lb" if m.group(1) == "bool" else "Li") + m.group(2) + "E"
    mangled_args = "I" + mangled_args + "E"
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = (short + mangled_args) in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
        paths[os.path.basename(m.group(1))] = m.group(1)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
        lines.append(cur)
if len(lines) != len(data):
    print("warning: %d SASS rows in the report, %d in the object (%s); is this the build that ran?" % (len(data), len(lines), short))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_ex = tot_s = 0
for r, l in zip(data, lines):
    ex, s, th = (int(r[ix[k]] or 0) for k in ("Instructions Executed", "# Samples", "Thread Instructions Executed"))
    a = agg[l]
    a[0] += ex
    a[1] += s
    a[2] += th
    tot_ex += ex
    tot_s += s
print("%s: %d warp-level instructions executed, %d stall samples" % (name.split("(")[0], tot_ex, tot_s))
print("%7s %7s %6s  %s" % ("exec %", "samp %", "lanes", "source line"))
src_cache = {}
for l, (ex, s, th) in sorted(agg.items(), key=lambda kv: -kv[1][0 if by_exec else 1])[:top]:
    f, n = l.rsplit(":", 1) if ":" in l else (l, "0")
    path = paths.get(f, os.path.join(os.path.dirname(obj), "..", "csrc", f))
    if f not in src_cache:
        src_cache[f] = open(path).read().splitlines() if os.path.exists(path) else []
    text = src_cache[f][int(n) - 1].strip()[:90] if 0 < int(n) <= len(src_cache[f]) else ""
    print("%6.1f%% %6.1f%% %6.1f  %-22s %s" % (100 * ex / max(tot_ex, 1), 100 * s / max(tot_s, 1), th / max(ex, 1), l, text))
