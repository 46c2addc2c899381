"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`): per-kernel headline metrics and, for one kernel, the SASS
hot spots.  Usage: python scripts/ncu_summary.py REPORT.ncu-rep [kernel-regex-for-source-page]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    print("== " + r[ik].split("(")[0])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("   %-85s %s %s" % (w, r[i], units[i]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + sys.argv[2]], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    isrc, isamp, iex, ith = (hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed",
                                                    "Thread Instructions Executed"))
    tot_ex = sum(int(r[iex]) for r in data) or 1
    tot_s = sum(int(r[isamp]) for r in data) or 1
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[iex])
        ops[o] += int(r[isamp])
    print("== SASS of %s: %d instructions, %d warp-level executions" % (sys.argv[2], len(data), tot_ex))
    for o, c in op.most_common(16):
        print("   %-8s executed %5.1f %%   stall samples %5.1f %%" % (o, 100 * c / tot_ex, 100 * ops[o] / tot_s))
    W = 200
    print("   windows of %d SASS instructions: first index, executed %%, samples %%, average active threads" % W)
    for i in range(0, len(data), W):
        blk = data[i:i + W]
        ex = sum(int(r[iex]) for r in blk)
        s = sum(int(r[isamp]) for r in blk)
        th = sum(int(r[ith]) for r in blk)
        if ex / tot_ex > 0.02 or s / tot_s > 0.02:
            print("   %6d  %5.1f %%  %5.1f %%  %5.1f" % (i, 100 * ex / tot_ex, 100 * s / tot_s, th / max(ex, 1)))
