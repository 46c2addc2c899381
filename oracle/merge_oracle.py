"""CPU restatement of the merge-filter step that follows a Fit-Hi-C run in the published protocol (reference
fithic/utils/CombineNearbyInteraction.py, driven by fithic/utils/merge-filter.sh).  TEST INFRASTRUCTURE ONLY: imported by
tests/ as the checker of fithic_b200/merge.py; nothing on the product path imports it.

Array level: the rows of a significances file (chr1, mid1, chr2, mid2, contactCount, p, q) in file order.  Pinned against the
unmodified reference (run here with networkx 3.6.1) by tests/golden/merge_*.npz (tests/golden/make_golden_merge.py).

What the reference does (CombineNearbyInteraction.py:236-727), per chromosome in `sort | uniq` order of column 1:
  * a node per distinct bin pair (bin = int(float(mid) + res / 2) / res, smaller bin first; the FIRST line of a repeated pair
    keeps its count / p / q, :293-307);
  * an edge between nodes whose two bins differ by <= 1 each (8-connectivity) or by <= 1 in total (4) (:313-333);
  * connected components, largest first, equal sizes in order of their first node (:346);
  * per component the bounding box, the sum of counts and the share of box cells that hold a node of ANY component
    (:362-406);
  * the representative loops: all nodes in (q, -count, bin1, bin2) order, greedily dropping a node when both bins lie within
    `Neigh` bins of a node already kept (-p 100, :585-712); the same but stopping at the top-K % q-value (0 < -p < 100,
    :458-577); or the single most significant node (-p 0, :421-456, an order-dependent partial comparison that walks the
    component in Python set order).
"""
import heapq


def custom_percent(lst, K, order=1):
    """CombineNearbyInteraction.py:38-52."""
    s = sorted(lst) if order == 1 else sorted(lst, reverse=True)
    index = int((len(lst) * K) / 100)
    if index <= 1:
        return max(s) if order == 1 else min(s)
    return s[index]


def chromosome_order(chr1_column):
    """`sort -k1,1 | uniq` of column 1 in the C locale (:195-205): byte order of the distinct names."""
    return sorted(set(chr1_column), key=lambda s: s.encode())


def bin_of(mid, res):
    """:297-299 (Python 3 true division: a float)."""
    return int(float(mid) + (res / 2)) / res


def _components(nodes, conn):
    """networkx 3.6.1 connected_components on the graph of :313-333: BFS from every unseen node in insertion order, the
    neighbours of a node in insertion order too (edges are added for i < j in lexicographic order); each component is the
    Python SET the BFS built, because -p 0 walks it in set order."""
    index = {k: i for i, k in enumerate(nodes)}
    if conn == 8:
        offs = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]
    elif conn == 4:
        offs = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    else:
        offs = []  # the reference adds no edge for any other value

    def adj(v):
        out = [index[(v[0] + a, v[1] + b)] for a, b in offs if (v[0] + a, v[1] + b) in index]
        return [nodes[i] for i in sorted(out)]

    seen_all = set()
    comps = []
    for v in nodes:
        if v in seen_all:
            continue
        seen = {v}
        nextlevel = [v]
        while nextlevel:
            thislevel, nextlevel = nextlevel, []
            for u in thislevel:
                for w in adj(u):
                    if w not in seen:
                        seen.add(w)
                        nextlevel.append(w)
        seen_all.update(seen)
        comps.append(seen)
    return sorted(comps, key=len, reverse=True)


def merge_chromosome(mid1, mid2, cc, pval, qval, res, conn=8, top_pct=100, neigh=2, sort_order=0):
    """The rows of one chromosome (intra lines only, file order) -> list of output rows
    (mid1, mid2, cc, p, q, span_low1, span_high1, span_low2, span_high2, sum_cc, share) in the reference's order."""
    thr = int(neigh) * res
    d = {}
    nodes = []
    for m1, m2, c, p, q in zip(mid1, mid2, cc, pval, qval):
        b1, b2 = bin_of(m1, res), bin_of(m2, res)
        key = (b1, b2) if b1 < b2 else (b2, b1)
        if key not in d:
            d[key] = (int(c), float(p), float(q))
            nodes.append(key)
    out = []
    for comp in _components(nodes, conn):
        members = list(comp)
        lo1, hi1 = int(min(x[0] for x in members)), int(max(x[0] for x in members))
        lo2, hi2 = int(min(x[1] for x in members)), int(max(x[1] for x in members))
        spans = ((lo1 - 1) * res, hi1 * res, (lo2 - 1) * res, hi2 * res)
        sum_cc = sum(d[x][0] for x in members)
        total = (hi1 - lo1 + 1) * (hi2 - lo2 + 1)
        have = sum(1 for a in range(lo1, hi1 + 1) for b in range(lo2, hi2 + 1) if (a, b) in d)
        share = (have * 1.0) / total
        reps = []
        if top_pct == 0:
            rep = members[0]
            for k in members[1:]:
                c, p, q = d[k]
                rc, rp, rq = d[rep]
                if sort_order == 0 and p < rp and q < rq:
                    rep = k
                elif sort_order == 1 and p > rp and q > rq:
                    rep = k
                elif p == rp and q == rq and c > rc:
                    rep = k
            reps = [rep]
        elif 0 < top_pct <= 100:  # any other value: none of the three branches runs, the component prints nothing
            heap = []
            for k in members:
                c, p, q = d[k]
                heapq.heappush(heap, [q if sort_order == 0 else -q, -c, k[0], k[1]])
            cut = None
            if top_pct < 100:
                cut = custom_percent([d[k][2] for k in members], top_pct, sort_order + 1)
            kept = []
            while heap:
                e = heapq.heappop(heap)
                if cut is not None and ((sort_order == 0 and e[0] > cut) or (sort_order == 1 and e[0] < cut)):
                    break  # (with -s 1 the heap holds -q, so this compares -q with a q: the reference's own behaviour)
                if kept and any(abs(a - e[2]) * res <= thr and abs(b - e[3]) * res <= thr for a, b in kept):
                    continue
                kept.append((e[2], e[3]))
            reps = kept
        for k in reps:
            c, p, q = d[k]
            low1, high1, low2, high2 = (k[0] - 1) * res, k[0] * res, (k[1] - 1) * res, k[1] * res
            out.append(((low1 + high1) / 2, (low2 + high2) / 2, c, p, q) + spans + (sum_cc, share))
    return out


HEADER = "\t".join(["chr1", "mid1", "chr2", "mid2", "CC", "p", "fdr", "bin1_low", "bin1_high", "bin2_low", "bin2_high", "sumCC",
                    "StrongConn"])


def merge_rows(chr1, mid1, chr2, mid2, cc, pval, qval, res, **kw):
    """All chromosomes -> the text of the output file (header without a trailing newline, every row preceded by one,
    :238, :455)."""
    text = [HEADER]
    for ch in chromosome_order(chr1):
        sel = [i for i in range(len(chr1)) if chr1[i] == ch and chr2[i] == ch]
        rows = merge_chromosome([mid1[i] for i in sel], [mid2[i] for i in sel], [cc[i] for i in sel], [pval[i] for i in sel],
                                [qval[i] for i in sel], res, **kw)
        for r in rows:
            text.append("\n" + "\t".join([ch, str(r[0]), ch, str(r[1])] + [str(v) for v in r[2:]]))
    return "".join(text)
