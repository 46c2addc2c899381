"""GPU, >= 2 devices: the real multi-GPU path (NCCL) against the oracle, one process per GPU under torchrun."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_pipeline_matches_oracle(lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_CHECK OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_two_gpu_cli_writes_the_single_gpu_file(lib, tmp_path):
    """`torchrun --nproc-per-node 2 -m fithic_b200 ...` and `python -m fithic_b200 ...` on the same gz inputs: the
    significances file (gunzipped) and the bin table are byte for byte the same, two spline passes, -x All."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import gzip

    from fithic_b200 import synth
    contacts, frags, biases, raw = synth.make_intra(120_000, 100000, seed=31, mean_count=4.0, with_bias=True,
                                                    inter_fraction=0.2)
    cpath, fpath, bpath = synth.write_inputs(str(tmp_path), contacts, frags, 100000, raw, biases, prefix="cli2")
    flags = ["-i", cpath, "-f", fpath, "-t", bpath, "-r", "100000", "-p", "2", "-x", "All", "-l", "cli2"]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    one = subprocess.run([sys.executable, "-m", "fithic_b200"] + flags + ["-o", str(tmp_path / "one")], capture_output=True,
                         text=True, timeout=600, cwd=ROOT, env=env)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519", "-m", "fithic_b200"] + flags +
                         ["-o", str(tmp_path / "two")], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    for name in ("cli2.spline_pass1.res100000.significances.txt.gz", "cli2.spline_pass2.res100000.significances.txt.gz"):
        with gzip.open(tmp_path / "one" / name, "rb") as f:
            a = f.read()
        with gzip.open(tmp_path / "two" / name, "rb") as f:
            b = f.read()
        assert len(a) > 1000 and a == b, name
    for name in ("cli2.fithic_pass2.res100000.txt",):
        assert (tmp_path / "one" / name).read_bytes() == (tmp_path / "two" / name).read_bytes()
    assert (tmp_path / "two" / "cli2.fithic_metrics.json").exists()
    assert not [f for f in os.listdir(tmp_path / "two") if ".part" in f]
