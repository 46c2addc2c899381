"""The ${lib}.fithic.log files the UNMODIFIED reference writes for the CLI golden cases (run in the build container, like
make_golden.py): tests/golden/<case>.fithic.log, with the output directory replaced by OUT.

    python tests/golden/make_golden_log.py
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fithic_b200 import synth  # noqa: E402
from oracle import ref_harness as R  # noqa: E402
from tests.golden.make_golden import CASES  # noqa: E402

for name in ("intra_bias_LU_p2", "all_bias", "inter_only_bias"):
    gen, flags = CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        contacts, frags, biases, raw = synth.make_intra(**gen)
        res = gen["res"]
        cpath, fpath, bpath = synth.write_inputs(tmp, contacts, frags, res, raw, biases, prefix=name)
        out = os.path.join(tmp, name + "_out")
        argv = ["-i", cpath, "-f", fpath, "-o", out, "-r", res, "-l", name] + flags
        if bpath:
            argv += ["-t", bpath]
        R.run_reference(argv)
        with open(os.path.join(out, name + ".fithic.log")) as f:
            text = f.read().replace(out, "OUT")
        with open(os.path.join(HERE, name + ".fithic.log"), "w") as f:
            f.write(text)
        print(name, len(text.splitlines()), "lines")
