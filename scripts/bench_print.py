"""Print the headline numbers of a bench.py JSON line (scripts/gpu_*.sh)."""
import json
import sys

d = json.loads([ln for ln in open(sys.argv[1]).read().strip().splitlines() if ln.startswith("{")][-1])
print("n%d ms/step %.3f value %.3g | e2e ms %.1f | digest %s | launches %s | host threads %s" % (
    d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("digest_line_p_q"), d["gpu_launches"],
    d.get("host_threads")))
print("   host", {k: round(v, 3) for k, v in d["host_ms_per_pass"].items()})
print("   kern", {k.replace("_kernel", ""): round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
if d.get("roofline"):
    r = d["roofline"]
    print("   roofline %s: %.0f GB/s of %.0f (%.2f), whole step %.2f" % (r["kernel"], r["achieved"], r["peak"], r["frac"],
                                                                        r["whole_step"]["frac"]))
    if r.get("fp64") and r["fp64"].get("kernels"):
        print("   fp64", {k: round(v["frac_of_burst_peak"], 3) for k, v in r["fp64"]["kernels"].items()},
              "peak", round(r["fp64"]["peak"]["burst_tflops"], 1))
for e in d.get("extra", []):
    if "failed" in e:
        print("   extra", e)
        continue
    print("   extra %s: ms/step %.3f value %.3g q<1 %d k4 %.3f digest %s" % (
        e["name"], e["ms_per_step"], e["value"], e["lines_with_q_below_1"], e["k4_ms_per_step"], e["digest_line_p_q"]))
