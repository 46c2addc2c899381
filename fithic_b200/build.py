"""Build libfithic_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The library is the product's only compute path; there is no CPU fallback.  `python -m fithic_b200.build` or
`__graft_entry__.build()` runs this; the resulting .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfithic_b200.so")
SOURCES = ["api.cu", "hist.cu", "host_bins.cu", "hoststage.cu", "spline.cu", "pvalue.cu", "pvalue_lists.cu", "bh.cu", "outlier.cu", "kr.cu", "merge.cu", "textio.cu", "peaks.cu", "comm.cu", "fragpairs.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "fithic_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into fithic_b200/libfithic_b200.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))] + [
        os.path.join(HERE, "..", "include", "fithic_b200.h"), os.path.abspath(__file__)]
    hdr_t = max(os.path.getmtime(h) for h in hdrs)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(path)):
            return obj, 0, ""
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = "".join(r[2] for r in results)
    bad = any(r[1] != 0 for r in results)
    if not bad:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [r[0] for r in results] + [
            "-lz", "-lpthread", "-o", LIB + ".tmp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        bad = res.returncode != 0
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if bad:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libfithic_b200.so")
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
