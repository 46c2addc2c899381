"""KR bias computation on the GPU: the `-t` file of Fit-Hi-C from raw contacts (reference fithic/utils/HiCKRy.py).

    python -m fithic_b200.hickry -i contacts.gz -f fragments.gz -o bias.gz [-x 0.05]

Same command line, same function names and the same output file (`chr<TAB>mid<TAB>bias`, -1 for the loci whose rows were
dropped) as the reference.  The contact lines are never turned into a matrix: they go to the GPU once as (row locus, column
locus, count) and every product with M + M^T is one pass of `fhc_kr_spmv` over them; the vector steps of the Knight-Ruiz
loop are the `fhc_kr_*` kernels (csrc/kr.cu).  The host keeps only the control flow of knightRuizAlg (HiCKRy.py:140-232):
a handful of scalars per iteration.  There is no CPU fallback.
"""
import argparse
import ctypes
import gzip
import sys
import time

import numpy as np
import torch

from . import _capi
from ._capi import check, dptr


def parse_args(arguments):
    parser = argparse.ArgumentParser(description="Check help flag")
    parser.add_argument("-i", "--interactions", help="Path to the interactions file to generate bias values", required=True,
                        type=str)
    parser.add_argument("-f", "--fragments", help="Path to the interactions file to generate bias values", required=True,
                        type=str)
    parser.add_argument("-o", "--output", help="Full path to output the generated bias file to", required=True, type=str)
    parser.add_argument("-x", "--percentOfSparseToRemove", help="Percent of diagonal to remove", required=False, type=float,
                        default=0.05)
    return parser.parse_args(arguments)


def loadfastfithicInteractions(interactionsFile, fragsFile):
    """HiCKRy.py:18-52 on arrays: loci are numbered in fragment-file order; returns ((x, y, z, n), revFrag) where a line
    adds z to M[x, y] and the matrix that is balanced is M + M^T."""
    import pandas as pd
    print("Creating sparse matrix...")
    startT = time.time()
    fr = pd.read_csv(fragsFile, sep=r"\s+", header=None, engine="c", usecols=[0, 2], names=["c", "m"],
                     dtype={"c": str, "m": np.int64}, compression="gzip")
    revFrag = list(zip(fr.c.tolist(), fr.m.tolist()))
    # the reference's dict keeps the LAST index of a repeated (chr, mid) (:29)
    key = pd.MultiIndex.from_arrays([fr.c, fr.m])
    idx = pd.Series(np.arange(len(fr), dtype=np.int64), index=key)
    idx = idx[~idx.index.duplicated(keep="last")]
    df = pd.read_csv(interactionsFile, sep=r"\s+", header=None, engine="c", names=["c1", "m1", "c2", "m2", "z"],
                     dtype={"c1": str, "c2": str, "m1": np.int64, "m2": np.int64, "z": np.float64}, compression="gzip",
                     float_precision="round_trip")
    x = idx.reindex(pd.MultiIndex.from_arrays([df.c1, df.m1])).to_numpy()
    y = idx.reindex(pd.MultiIndex.from_arrays([df.c2, df.m2])).to_numpy()
    if np.isnan(x).any() or np.isnan(y).any():
        raise KeyError("a contact names a locus that is not in the fragments file")  # the reference raises KeyError too
    endT = time.time()
    print("Sparse matrix creation took %s seconds" % (endT - startT))
    return (x.astype(np.int32), y.astype(np.int32), df.z.to_numpy(np.float64), len(fr)), revFrag


class KRDevice:
    """The contact lines and the work vectors of one balancing run on the GPU."""

    def __init__(self, x, y, z, n, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fithic_b200.hickry needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _capi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_all = int(n)
        self.rows = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(self.device)
        self.cols = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32)).to(self.device)
        self.vals = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64)).to(self.device)
        self.nnz = self.rows.numel()
        self.P = int(self.lib.fhc_kr_partials())
        self.partial = torch.zeros(2 * self.P, dtype=torch.float64, device=self.device)
        self.set_kept(np.arange(self.n_all))

    @classmethod
    def from_device(cls, rows, cols, vals, n):
        """Adopt device tensors (int32, int32, float64) instead of copying host arrays."""
        self = cls.__new__(cls)
        self.lib = _capi.load()
        self.device = rows.device
        self.n_all = int(n)
        self.rows, self.cols, self.vals = rows.contiguous(), cols.contiguous(), vals.contiguous()
        self.nnz = self.rows.numel()
        self.P = int(self.lib.fhc_kr_partials())
        self.partial = torch.zeros(2 * self.P, dtype=torch.float64, device=self.device)
        self.set_kept(np.arange(self.n_all))
        return self

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_kept(self, kept):
        """Loci that stay in the matrix (ascending original indices); everything else is dropped from rows and columns."""
        remap = np.full(self.n_all, -1, dtype=np.int32)
        remap[np.asarray(kept, dtype=np.int64)] = np.arange(len(kept), dtype=np.int32)
        self.remap = torch.from_numpy(remap).to(self.device)
        self.n = int(len(kept))

    def vec(self, fill=None):
        t = torch.empty(max(self.n, 1), dtype=torch.float64, device=self.device)[:self.n]
        if fill is not None:
            t.fill_(fill)
        return t

    def spmv(self, x, y):
        check(self.lib.fhc_kr_spmv(dptr(self.rows), dptr(self.cols), dptr(self.vals), self.nnz, dptr(self.remap), dptr(x),
                                   dptr(y), self.n, self._stream()))
        return y

    def _sum(self):
        return float(self.partial[:self.P].cpu().numpy().sum())


def row_sums(dev):
    """mtx.sum(axis=0) of the full matrix (HiCKRy.py:79): one product with the all-ones vector."""
    ones = dev.vec(1.0)
    out = dev.vec()
    dev.spmv(ones, out)
    return out.cpu().numpy()


def removeZeroDiagonalCSR(rowSums, perc):
    """HiCKRy.py:76-96 on the row sums: the int(perc * size)-th smallest sum is the threshold, every locus with a sum
    <= threshold is removed.  Returns the sorted list of removed loci."""
    rs = np.asarray(rowSums, dtype=np.float64).reshape(-1)
    size = len(rs)
    rem = int(perc * size)
    print("Removing %s percent of most sparse bins" % (perc))
    print("... corresponds to %s total rows" % (rem))
    order = np.argsort(rs, kind="stable")
    valToRemove = rs[order[rem]]
    print("... corresponds to all bins with less than or equal to %s total interactions" % valToRemove)
    return np.nonzero(rs <= valToRemove)[0].tolist()


def knightRuizAlg(dev, tol=1e-6, f1=False):
    """knightRuizAlg (HiCKRy.py:140-232) with the reference's control flow on the host and every vector on the GPU.
    Returns [x (device tensor), outer iterations, inner iterations of the last outer one]."""
    lib, n, st = dev.lib, dev.n, dev._stream
    P, part = dev.P, dev.partial
    Delta, delta, g = 3, 0.1, 0.9
    etamax = eta = 0.1
    stop_tol = tol * 0.5
    x = dev.vec(1.0)
    rt = tol ** 2.0
    Ax, v, rk, Z, p, w, xp, y = (dev.vec() for _ in range(8))
    dev.spmv(x, Ax)
    check(lib.fhc_kr_residual(dptr(x), dptr(Ax), dptr(v), dptr(rk), n, dptr(part), st()))
    rho_km1 = dev._sum()
    rho_km2 = rho_km1
    rout = rold = rho_km1
    MVP = 0
    i = 0
    k = 0
    while rout > rt:  # outer iteration
        i += 1
        if i > 30:
            break
        k = 0
        y.fill_(1.0)
        innertol = max(eta ** 2.0 * rout, rt)
        while rho_km1 > innertol:  # inner iteration by CG
            k += 1
            if k == 1:
                check(lib.fhc_kr_first(dptr(rk), dptr(v), dptr(Z), dptr(p), n, dptr(part), st()))
                rho_km1 = dev._sum()
                check(lib.fhc_kr_direction(dptr(Z), 0.0, 1, dptr(p), dptr(x), dptr(xp), n, st()))
            else:
                beta = rho_km1 / rho_km2
                check(lib.fhc_kr_direction(dptr(Z), float(beta), 0, dptr(p), dptr(x), dptr(xp), n, st()))
            if k > 10:
                break
            dev.spmv(xp, Ax)  # A.dot(x * p)
            check(lib.fhc_kr_w(dptr(x), dptr(Ax), dptr(v), dptr(p), dptr(w), n, dptr(part), st()))
            alpha = rho_km1 / dev._sum()
            check(lib.fhc_kr_ynew_minmax(dptr(y), float(alpha), dptr(p), n, dptr(part), st()))
            mm = part.cpu().numpy()
            ymin, ymax = float(mm[:P].min()), float(-mm[P:].min())
            if ymin <= delta:
                if delta == 0:
                    break
                check(lib.fhc_kr_gamma(dptr(y), float(alpha), dptr(p), float(delta), 0, n, dptr(part), st()))
                gamma = float(part[:P].cpu().numpy().min())
                check(lib.fhc_kr_axpy(dptr(y), float(gamma), float(alpha), dptr(p), n, st()))
                break
            if ymax >= Delta:
                check(lib.fhc_kr_gamma(dptr(y), float(alpha), dptr(p), float(Delta), 1, n, dptr(part), st()))
                gamma = float(part[:P].cpu().numpy().min())
                check(lib.fhc_kr_axpy(dptr(y), float(gamma), float(alpha), dptr(p), n, st()))
                break
            rho_km2 = rho_km1
            check(lib.fhc_kr_update(dptr(y), float(alpha), dptr(p), dptr(rk), dptr(w), dptr(v), dptr(Z), n, dptr(part), st()))
            rho_km1 = dev._sum()
        check(lib.fhc_kr_mul(dptr(x), dptr(y), dptr(x), n, st()))  # x *= y
        dev.spmv(x, Ax)
        check(lib.fhc_kr_residual(dptr(x), dptr(Ax), dptr(v), dptr(rk), n, dptr(part), st()))
        rho_km1 = dev._sum()
        rout = rho_km1
        MVP += k + 1
        rat = rout / rold
        rold = rout
        res_norm = rout ** 0.5
        eta_o = eta
        eta = g * rat
        if g * eta_o ** 2.0 > 0.1:
            eta = max(eta, g * eta_o ** 2.0)
        eta = max(min(eta, etamax), stop_tol / res_norm)
        if f1:
            print("%03i %06i %03.3f %e %e" % (i, k, res_norm, rt, rout))
    if f1:
        print("Matrix - vector products = %06i" % MVP)
    return [x, i, k]


def computeBiasVector(x):
    """HiCKRy.py:98-104 (n values on the host)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    one = np.ones((x.shape[0], 1))
    x = one / x
    sums = np.sum(x)
    avg = (1.0 * sums) / x.shape[0]
    return np.divide(x, avg)


def addZeroBiases(lst, vctr):
    """HiCKRy.py:106-109: -1 at every removed locus."""
    n = len(vctr) + len(lst)
    out = np.full((n, 1), -1.0)
    keep = np.ones(n, dtype=bool)
    keep[np.asarray(lst, dtype=np.int64)] = False
    out[keep] = np.asarray(vctr, dtype=np.float64).reshape(-1, 1)
    return out


def returnBias(rawMatrix, perc, device=None):
    """HiCKRy.py:53-74.  rawMatrix = (x, y, z, n) from loadfastfithicInteractions."""
    x, y, z, n = rawMatrix
    dev = KRDevice(x, y, z, n, device)
    removed = removeZeroDiagonalCSR(row_sums(dev), perc)
    print("Sparse rows removed")
    print("Initial matrix size: %s rows and %s columns" % (n, n))
    keep = np.ones(n, dtype=bool)
    keep[removed] = False
    dev.set_kept(np.nonzero(keep)[0])
    print("New matrix size: %s rows and %s columns" % (dev.n, dev.n))
    print("Normalizing with KR Algorithm")
    result = knightRuizAlg(dev)
    bias = computeBiasVector(result[0].cpu().numpy())
    returnBias.last = dict(removed=removed, outer=result[1], inner=result[2])
    return addZeroBiases(removed, bias)


def checkBias(biasvec):
    """HiCKRy.py:234-250."""
    b = np.asarray(biasvec)
    std, mean, median = np.std(b), np.mean(b), np.median(b)
    if (mean < 0.5 or mean > 2) or (median < 0.5 or median > 2):
        which = "mean" if (mean < 0.5 or mean > 2) else "median"
        print("WARNING... Bias vector has a %s outside of typical range (0.5, 2)." % which)
        print("Consider running with a larger -x option if problems occur")
        print("Mean\t%s" % mean)
        print("Median\t%s" % median)
        print("Std. Dev.\t%s" % std)


def outputBias(biasCol, revFrag, outputFilePath):
    """HiCKRy.py:252-262: chr<TAB>mid<TAB>bias, the value printed like the reference's `%s` of a numpy float64."""
    with gzip.open(outputFilePath, "wt") as biasFile:
        for (chrom, mid), value in zip(revFrag, np.asarray(biasCol, dtype=np.float64).reshape(-1)):
            biasFile.write("%s\t%s\t%s\n" % (chrom, mid, np.float64(value)))


def main(argv=None):
    args = parse_args(sys.argv[1:] if argv is None else argv)
    matrix, revFrag = loadfastfithicInteractions(args.interactions, args.fragments)
    bias = returnBias(matrix, args.percentOfSparseToRemove)
    checkBias(bias)
    outputBias(bias, revFrag, args.output)


if __name__ == "__main__":
    main()
