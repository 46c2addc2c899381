"""Merge-filter step after a Fit-Hi-C run: the reference's fithic/utils/CombineNearbyInteraction.py with
the same flags, the same output file byte for byte, and the connected components, box statistics and greedy choice of
representative loops computed on the GPU (csrc/merge.cu).  `merge_filter` is fithic/utils/merge-filter.sh.

What stays on the host: reading the text, the order of the components in the output (a lexsort over the components), the
formatting of the rows with Python's own str() of ints and floats, and the single-representative mode `-p 0`, whose result
depends on the order in which Python iterates the SET networkx builds for a component (CombineNearbyInteraction.py:355,
:423-444) and is therefore replayed with the same set operations on the GPU's components.

Not reproduced: the `Temp_<chr>_Dump.bed` / `<out>_chrName.bed` scratch files and the per-component debug prints.  Mid
points must lie on the bin grid (int(mid + res / 2) divisible by res), as they do in every Fit-Hi-C output with a fixed
bin size; the reference would carry fractional bin numbers through its comparisons, which is refused here.
"""
import argparse
import gzip
import os
import sys

import numpy as np

from . import _capi

HEADER = "\t".join(["chr1", "mid1", "chr2", "mid2", "CC", "p", "fdr", "bin1_low", "bin1_high", "bin2_low", "bin2_high", "sumCC",
                    "StrongConn"])
MAX_BIN = (1 << 24) - 2


def parse_args(args):
    """The flags, destinations, types and defaults of CombineNearbyInteraction.py:83-109."""
    parser = argparse.ArgumentParser(description="Merge nearby significant Fit-Hi-C interactions (connected components) on the GPU")
    parser.add_argument("-i", "--InpFile", required=True, help="significances file (.gz or plain text)")
    parser.add_argument("-H", "--headerInp", dest="headerInp", type=int, default=1, help="1: the input has a header line (default)")
    parser.add_argument("-o", "--OutFile", required=True, help="merged interactions, gzipped")
    parser.add_argument("-r", "--resolution", required=True, help="bin size of the Fit-Hi-C run")
    parser.add_argument("-c", "--conn", dest="connectivity_rule", type=int, default=8, required=False,
                        help="8 (default) or 4: which neighbouring bin pairs are connected")
    parser.add_argument("-p", "--percent", dest="TopPctElem", type=int, default=100,
                        help="100 (default): every loop of a component is a candidate; 0: only the most significant one; "
                             "x in between: the top x %% by q-value")
    parser.add_argument("-n", "--Neigh", dest="NeighborHoodBin", type=int, default=2,
                        help="a candidate is dropped when both its bins lie within this many bins of a loop already kept")
    parser.add_argument("-s", "--order", dest="SortOrder", type=int, default=0,
                        help="0 (default): smaller significance values are better; 1: larger are better")
    return parser.parse_args(args)


def read_rows(path, header=1, fdr=None):
    """The first seven columns of a significances file (whitespace separated, like awk and str.split see them).
    fdr: keep the rows with q <= fdr (the awk line of merge-filter.sh:22)."""
    import pandas as pd
    opener = gzip.open if path.endswith(".gz") else open
    try:
        with opener(path, "rt") as f:
            df = pd.read_csv(f, sep=r"\s+", header=None, skiprows=1 if header == 1 else 0, usecols=range(7),
                             names=["c1", "m1", "c2", "m2", "cc", "p", "q"], engine="c", float_precision="round_trip",
                             dtype={"c1": str, "m1": np.float64, "c2": str, "m2": np.float64, "cc": np.int64, "p": np.float64,
                                    "q": np.float64})
    except pd.errors.EmptyDataError:
        df = pd.DataFrame({k: [] for k in ["c1", "m1", "c2", "m2", "cc", "p", "q"]})
    if fdr is not None:
        df = df[df.q.to_numpy(np.float64) <= float(fdr)]
    return dict(chr1=df.c1.to_numpy(object), chr2=df.c2.to_numpy(object), mid1=df.m1.to_numpy(np.float64),
                mid2=df.m2.to_numpy(np.float64), cc=df.cc.to_numpy(np.int64), p=df.p.to_numpy(np.float64),
                q=df.q.to_numpy(np.float64))


def chromosome_order(chr1):
    """`sort -k1,1 | uniq` of column 1 (CombineNearbyInteraction.py:195-205), C locale."""
    return sorted(set(chr1.tolist()), key=lambda s: s.encode())


def bins_of(mid, res):
    """int(float(mid) + res / 2) / res (:297-299) for mid points on the bin grid; anything else is refused."""
    n = np.trunc(np.asarray(mid, dtype=np.float64) + res / 2).astype(np.int64)
    if np.any(n % res != 0):
        bad = np.asarray(mid)[np.nonzero(n % res != 0)[0][0]]
        raise ValueError("mid point %r is not on the %d bp bin grid (fractional bin numbers are not supported)" % (bad, res))
    b = n // res
    if len(b) and (b.min() < 1 or b.max() > MAX_BIN):
        raise ValueError("bin numbers must lie in [1, %d]" % MAX_BIN)
    return b


def components_device(chr_rank, b1, b2, cc, q, conn, top_pct, neigh, sort_order, device=None):
    """The two library calls on device copies of the per-line arrays -> dict of host arrays (see include/fithic_b200.h)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("fithic_b200.merge needs a CUDA device (there is no CPU fallback)")
    lib = _capi.load()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    n = len(b1)
    with torch.cuda.device(dev):
        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        d_chr, d_b1, d_b2 = up(chr_rank, np.int32), up(b1, np.int32), up(b2, np.int32)
        d_cc, d_q = up(cc, np.int64), up(q, np.float64)
        m = max(n, 1)
        keys = torch.empty(m, dtype=torch.int64, device=dev)
        order = torch.empty(m, dtype=torch.int32, device=dev)
        label = torch.empty(m, dtype=torch.int32, device=dev)
        size = torch.empty(m, dtype=torch.int32, device=dev)
        first_line = torch.empty(m, dtype=torch.int32, device=dev)
        box = torch.empty(4 * m, dtype=torch.int32, device=dev)
        sum_cc = torch.empty(m, dtype=torch.int64, device=dev)
        have = torch.empty(m, dtype=torch.int64, device=dev)
        ranked = torch.empty(m, dtype=torch.int32, device=dev)
        keep = torch.zeros(m, dtype=torch.uint8, device=dev)
        ws_bytes = lib.fhc_merge_workspace_bytes(n)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(lib.fhc_merge_components(d_chr.data_ptr(), d_b1.data_ptr(), d_b2.data_ptr(), d_cc.data_ptr(), n, int(conn),
                                             keys.data_ptr(), order.data_ptr(), label.data_ptr(), size.data_ptr(),
                                             first_line.data_ptr(), box.data_ptr(), sum_cc.data_ptr(), have.data_ptr(),
                                             ws.data_ptr(), ws_bytes, stream))
        select = 0 < top_pct <= 100
        if select:
            _capi.check(lib.fhc_merge_select(keys.data_ptr(), order.data_ptr(), label.data_ptr(), size.data_ptr(),
                                             d_cc.data_ptr(), d_q.data_ptr(), n, int(top_pct), int(neigh), int(sort_order),
                                             ranked.data_ptr(), keep.data_ptr(), ws.data_ptr(), ws_bytes, stream))
        torch.cuda.synchronize(dev)
        out = dict(keys=keys[:n].cpu().numpy().view(np.uint64), order=order[:n].cpu().numpy().view(np.uint32),
                   label=label[:n].cpu().numpy(), size=size[:n].cpu().numpy(),
                   first_line=first_line[:n].cpu().numpy().view(np.uint32), box=box[:4 * n].cpu().numpy().reshape(-1, 4),
                   sum_cc=sum_cc[:n].cpu().numpy(), have=have[:n].cpu().numpy())
        if select:
            out["ranked"] = ranked[:n].cpu().numpy().view(np.uint32)
            out["keep"] = keep[:n].cpu().numpy()
    return out


def _set_order_members(comp_nodes, conn):
    """The members of one component in the order `list(component)` has in the reference (:355): networkx builds the
    component as a Python set by BFS from its first node, neighbours in node insertion order."""
    index = {k: i for i, k in enumerate(comp_nodes)}  # comp_nodes: (bin1, bin2) float tuples in insertion (line) order
    if conn == 8:
        offs = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]
    elif conn == 4:
        offs = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    else:
        offs = []
    seen = {comp_nodes[0]}
    nextlevel = [comp_nodes[0]]
    while nextlevel:
        thislevel, nextlevel = nextlevel, []
        for v in thislevel:
            near = sorted(index[(v[0] + a, v[1] + b)] for a, b in offs if (v[0] + a, v[1] + b) in index)
            for i in near:
                w = comp_nodes[i]
                if w not in seen:
                    seen.add(w)
                    nextlevel.append(w)
    return list(seen)


def _single_representative(members, value, sort_order):
    """:423-444 -- the partial comparison that keeps the first member unless a later one is better in BOTH p and q."""
    rep = members[0]
    for k in members[1:]:
        c, p, q = value[k]
        rc, rp, rq = value[rep]
        if sort_order == 0 and p < rp and q < rq:
            rep = k
        elif sort_order == 1 and p > rp and q > rq:
            rep = k
        elif p == rp and q == rq and c > rc:
            rep = k
    return rep


def merge_rows(rows, res, conn=8, top_pct=100, neigh=2, sort_order=0, log=None, components=None):
    """rows: dict from read_rows -> the text of the reference's output file.  components: the routine that produces the
    component arrays (tests pass the library's serial host drivers here; the default and only product path is the GPU)."""
    res = int(res)
    if sort_order not in (0, 1):
        raise ValueError("-s / --order must be 0 or 1")
    chr1, chr2 = rows["chr1"], rows["chr2"]
    names = chromosome_order(chr1) if len(chr1) else []
    if log:
        log("List of chromosomes considered:  %s" % str(names))
    intra = np.nonzero(chr1 == chr2)[0] if len(chr1) else np.zeros(0, np.int64)
    if len(intra) == 0:
        return HEADER
    if len(names) > 65535:
        raise ValueError("more than 65535 chromosome names")
    rank = {c: i for i, c in enumerate(names)}
    chr_rank = np.fromiter((rank[c] for c in chr1[intra].tolist()), dtype=np.int32, count=len(intra))
    ba, bb = bins_of(rows["mid1"][intra], res), bins_of(rows["mid2"][intra], res)
    b1, b2 = np.minimum(ba, bb), np.maximum(ba, bb)
    cc, pv, qv = rows["cc"][intra], rows["p"][intra], rows["q"][intra]
    if len(cc) and (cc.min() < 0 or cc.max() >= 1 << 31):
        raise ValueError("contact counts must lie in [0, 2^31)")
    if np.isnan(qv).any():
        raise ValueError("NaN q-values cannot be ordered (filter the file by q first, as merge-filter.sh does)")
    run = components if components is not None else components_device
    g = run(chr_rank, b1, b2, cc, qv, conn, top_pct, neigh, sort_order)
    keys, order, label = g["keys"], g["order"], g["label"]
    n = len(keys)
    roots = np.nonzero(label == np.arange(n, dtype=np.int64))[0]
    root_chr = (keys[roots] >> np.uint64(48)).astype(np.int64)
    by = np.lexsort((g["first_line"][roots].astype(np.int64), -g["size"][roots].astype(np.int64), root_chr))
    roots = roots[by]  # the reference's order: chromosome, larger component first, then the one whose first node came first
    if log:
        for ci, name in enumerate(names):
            sel = root_chr[by] == ci
            if sel.any():
                log("Processing the chromosome:  %s" % name)
                log("No of nodes of G:  %d" % int(g["size"][roots[sel]].sum()))
                log("Number of connected components of G:  %d" % int(sel.sum()))
    comp_rank = np.full(n, -1, dtype=np.int64)
    comp_rank[roots] = np.arange(len(roots))
    eb1 = ((keys >> np.uint64(24)) & np.uint64(0xffffff)).astype(np.int64)
    eb2 = (keys & np.uint64(0xffffff)).astype(np.int64)

    if top_pct == 0:
        # nodes of every component in insertion (line) order, then Python's own set order
        node_e = np.nonzero(label >= 0)[0]
        node_e = node_e[np.lexsort((order[node_e].astype(np.int64), comp_rank[label[node_e]]))]
        cuts = np.searchsorted(comp_rank[label[node_e]], np.arange(len(roots) + 1))
        picked = []
        for c in range(len(roots)):
            es = node_e[cuts[c]:cuts[c + 1]]
            nodes = [(float(a), float(b)) for a, b in zip(eb1[es].tolist(), eb2[es].tolist())]
            value = {k: (int(cc[order[e]]), float(pv[order[e]]), float(qv[order[e]])) for k, e in zip(nodes, es.tolist())}
            entry = dict(zip(nodes, es.tolist()))
            picked.append(entry[_single_representative(_set_order_members(nodes, conn), value, sort_order)])
        picked = np.asarray(picked, dtype=np.int64)
    elif 0 < top_pct <= 100:
        w = np.nonzero(g["keep"])[0]
        e = g["ranked"][w].astype(np.int64)
        picked = e[np.lexsort((w, comp_rank[label[e]]))]
    else:
        picked = np.zeros(0, dtype=np.int64)  # none of the three branches of the reference runs

    text = [HEADER]
    box, sum_cc, have = g["box"], g["sum_cc"], g["have"]
    for e in picked.tolist():
        r = int(label[e])
        line = int(order[e])
        name = names[int(keys[e] >> np.uint64(48))]
        k0, k1 = float(eb1[e]), float(eb2[e])
        mid_a = (((k0 - 1) * res) + (k0 * res)) / 2
        mid_b = (((k1 - 1) * res) + (k1 * res)) / 2
        lo1, hi1, lo2, hi2 = (int(v) for v in box[r])
        total = (hi1 - lo1 + 1) * (hi2 - lo2 + 1)
        share = (int(have[r]) * 1.0) / total
        text.append("\n" + "\t".join([name, str(mid_a), name, str(mid_b), str(int(cc[line])), str(float(pv[line])),
                                     str(float(qv[line])), str((lo1 - 1) * res), str(hi1 * res), str((lo2 - 1) * res),
                                     str(hi2 * res), str(int(sum_cc[r])), str(share)]))
    return "".join(text)


def combine_nearby_interactions(InpFile, OutFile, resolution, headerInp=1, connectivity_rule=8, TopPctElem=100,
                                NeighborHoodBin=2, SortOrder=0, fdr=None, log=print, components=None):
    """CombineNearbyInteraction.py main() (:111-727) on files."""
    bin_size = int(resolution)
    if log:
        for label, v in (("bin_size", bin_size), ("headerInp", int(headerInp)), ("connectivity_rule", int(connectivity_rule)),
                         ("TopPctElem", int(TopPctElem)), ("NeighborHoodBinThr", int(NeighborHoodBin) * bin_size), ("QValCol", 7),
                         ("PValCol", 6), ("SortOrder", int(SortOrder))):
            log("\n *** %s:  %s" % (label, v))
    out_dir = os.path.dirname(os.path.realpath(OutFile))
    os.makedirs(out_dir, exist_ok=True)
    rows = read_rows(InpFile, int(headerInp), fdr)
    text = merge_rows(rows, bin_size, int(connectivity_rule), int(TopPctElem), int(NeighborHoodBin), int(SortOrder), log, components)
    with gzip.open(OutFile, "wt") as f:
        f.write(text)
    if log:
        log("End of merging filtering loops !!! ")
    return text.count("\n")


def merge_filter(inputFile, resolution, outputFile, fdr, log=None, components=None):
    """fithic/utils/merge-filter.sh: drop the header line, keep the rows with q <= fdr, merge with the default options
    (the intermediate fithic_subset.gz of the script is not written)."""
    return combine_nearby_interactions(inputFile, outputFile, resolution, headerInp=1, fdr=fdr, log=log, components=components)


def main(argv=None):
    """`python -m fithic_b200.merge -i ... -o ... -r ...` is CombineNearbyInteraction.py; with positional arguments
    `inputFile resolution outputFile fdr [utilityFolder]` it is merge-filter.sh (the last argument is accepted and ignored)."""
    argv = sys.argv[1:] if argv is None else list(argv)
    if argv and not argv[0].startswith("-"):
        if len(argv) not in (4, 5):
            sys.exit("usage: python -m fithic_b200.merge inputFile resolution outputFile fdr [utilityFolder]")
        merge_filter(argv[0], int(argv[1]), argv[2], float(argv[3]), log=print)
        return
    o = parse_args(argv)
    combine_nearby_interactions(o.InpFile, o.OutFile, o.resolution, o.headerInp, o.connectivity_rule, o.TopPctElem,
                                o.NeighborHoodBin, o.SortOrder)


if __name__ == "__main__":
    main()
