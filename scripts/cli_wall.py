"""End-to-end wall time of the command line on a synthetic whole-genome contact file (text in, gz text out): what a user of
the reference's CLI sees.  Usage: python scripts/cli_wall.py [lines] [resolution]   (one GPU; the file is written first and
not timed).  Prints the CLI's own metrics sidecar (${lib}.fithic_metrics.json) and the wall clock around `main`."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fithic_b200 import fithic as cli  # noqa: E402
from fithic_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
res = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
with tempfile.TemporaryDirectory() as tmp:
    t0 = time.time()
    contacts, frags, biases, raw = synth.make_intra(n, res, seed=77, mean_count=4.0, with_bias=True)
    cpath, fpath, bpath = synth.write_inputs(tmp, contacts, frags, res, raw, biases, prefix="wall")
    print("input written in %.1f s: %d lines, %.1f MB gz" % (time.time() - t0, n, os.path.getsize(cpath) / 1e6))
    out = os.path.join(tmp, "out")
    argv = ["-i", cpath, "-f", fpath, "-o", out, "-r", str(res), "-t", bpath, "-l", "wall", "-p", "2"]
    t0 = time.time()
    cli.main(argv)
    wall = time.time() - t0
    with open(os.path.join(out, "wall.fithic_metrics.json")) as f:
        m = json.load(f)
    sig = os.path.join(out, "wall.spline_pass2.res%d.significances.txt.gz" % res)
    print("CLI wall %.2f s for %d lines and 2 spline passes = %.2e lines/s (output %.1f MB gz)" % (wall, n, n / wall,
                                                                                             os.path.getsize(sig) / 1e6))
    print(json.dumps(m, indent=1))
