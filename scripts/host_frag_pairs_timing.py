"""Times fhc_host_frag_pairs with 1/2/4/8 threads on a 5 kb whole-genome bin table (host only)."""
import sys, time, os, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from fithic_b200 import synth, _capi
from fithic_b200 import engine as E
lib=_capi.load()
res=5000
contacts, frags, biases, _ = synth.make_intra(3_000_000, res, seed=1004, mean_count=3.0, with_bias=True)
d=np.abs(contacts.mid1.astype(np.int64)-contacts.mid2)//res
D=int(max(d.max(), frags.max_mid.max()//res))+2
hist=np.bincount(d, weights=contacts.cnt, minlength=D).astype(np.int64)*100
N=int(hist.sum())
st=E.Settings(resolution=res,noOfBins=100)
seen=np.nonzero(hist)[0]
bins0 = E.make_bins(lib, (seen*res).astype(np.int64), hist[seen].astype(np.int64), 100, N)
import copy
for nt in (1,2,4,8):
    os.environ["FHC_HOST_THREADS"]=str(nt)
    ts=[]
    for i in range(30):
        b=copy.deepcopy(bins0)
        t=time.perf_counter(); fp=E.frag_pairs(lib, frags, st, b, None); ts.append(time.perf_counter()-t)
    print(nt, "threads: %.3f ms (min %.3f)" % (np.mean(ts[5:])*1e3, np.min(ts)*1e3), b["sumdist"][-1], b["pairs"][-3:])
w=[(bins0["ub"][i]-bins0["lb"][i])//res for i in range(bins0["n"])]
print("bin widths (steps): last 6", w[-6:], "total", sum(w))
from fithic_b200._capi import dptr, check
order = sorted(range(len(frags.chroms)), key=lambda i: frags.chroms[i])
order = [i for i in order if frags.n_mappable[i] > 0]
chr_n = np.ascontiguousarray(frags.n_mappable[order], dtype=np.int64)
chr_mm = np.ascontiguousarray(frags.max_mid[order], dtype=np.int64)
nb = bins0["n"]
for nt in (1,4):
    os.environ["FHC_HOST_THREADS"]=str(nt)
    ts=[]
    for i in range(30):
        pairs = np.zeros(nb, dtype=np.int64); sumdist = np.zeros(nb); totals = np.zeros(4, dtype=np.int64)
        t=time.perf_counter()
        check(lib.fhc_host_frag_pairs(dptr(chr_n), dptr(chr_mm), len(order), res, st.L, st.U, dptr(bins0["lb"]), dptr(bins0["ub"]), nb, dptr(pairs), dptr(sumdist), dptr(totals)))
        ts.append(time.perf_counter()-t)
    print(nt, "C call only: %.3f ms (min %.3f)" % (np.mean(ts[5:])*1e3, np.min(ts)*1e3))
