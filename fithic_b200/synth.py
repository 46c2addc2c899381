"""Deterministic synthetic Hi-C inputs of the shapes BASELINE.json names.

Genome = hg38 chr1-22,X,Y; fixed-size fragments laid out like the reference's
utils/createFitHiCFragments-fixedsize.py:57-76 (start = k*res, mid = start + res/2, hits = 1, last partial bin kept).
Contacts: locus pairs (i <= j) on one chromosome with P(j - i = k) ~ 1/(k+1); count = 1 + Poisson(lam0 (k+1)^-1.08 b_i b_j).
`numpy` generator for tests (small n, bit reproducible), `torch` generator on the GPU for the 300 M pair bench input.
"""
import numpy as np

from .engine import Biases, Contacts, Fragments

HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
        ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
        ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


def genome(chroms=None):
    g = HG38 if chroms is None else [c for c in HG38 if c[0] in chroms]
    return [c[0] for c in g], np.array([c[1] for c in g], dtype=np.int64)


def n_bins(sizes, res):
    return (sizes + res - 1) // res


def fragments_for(names, sizes, res):
    nb = n_bins(sizes, res)
    return Fragments(list(names), nb.astype(np.int64), ((nb - 1) * res + res // 2).astype(np.int64))


def make_biases(names, sizes, res, rng, frac_out=0.03, frac_nan=0.005):
    """LogNormal(0, 0.25) rescaled to mean 1; some loci pushed outside [0.5, 2] and some NaN (-> -1 after read_biases)."""
    nb = n_bins(sizes, res)
    off = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum(nb, out=off[1:])
    tot = int(off[-1])
    b = rng.lognormal(0.0, 0.25, tot)
    b /= b.mean()
    u = rng.random(tot)
    b[u < frac_out / 2] = 0.3
    b[(u >= frac_out / 2) & (u < frac_out)] = 2.5
    b[(u >= frac_out) & (u < frac_out + frac_nan)] = np.nan
    raw = b.copy()
    bad = (b < 0.5) | np.isnan(b) | (b > 2.0)
    vals = np.where(bad, -1.0, b)
    mids = np.concatenate([np.arange(n, dtype=np.int64) * res + res // 2 for n in nb]).astype(np.int32)
    return Biases(vals.astype(np.float64), mids, off), raw


def make_intra(n_pairs, res, seed, chroms=None, mean_count=3.0, with_bias=False, inter_fraction=0.0):
    """numpy generator (tests / small configs).  Returns (Contacts, Fragments, Biases or None, raw bias values)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    names, sizes = genome(chroms)
    nb = n_bins(sizes, res)
    frags = fragments_for(names, sizes, res)
    biases, raw = (make_biases(names, sizes, res, rng) if with_bias else (None, None))
    w = sizes / sizes.sum()
    per = rng.multinomial(n_pairs, w)
    m1s, m2s, cs, chs = [], [], [], []
    # lam0 such that the mean count is ~mean_count under P(k) ~ 1/(k+1)
    kmax = int(nb.max())
    ks = np.arange(kmax)
    pk = 1.0 / (ks + 1.0)
    pk /= pk.sum()
    lam0 = (mean_count - 1.0) / float((pk * (ks + 1.0) ** -1.08).sum())
    for ci, n in enumerate(per):
        if n == 0:
            continue
        nbc = int(nb[ci])
        u = rng.random(n)
        k = np.minimum(np.floor((nbc + 1.0) ** u).astype(np.int64) - 1, nbc - 1)
        i = np.floor(rng.random(n) * (nbc - k)).astype(np.int64)
        j = i + k
        lam = lam0 * (k + 1.0) ** -1.08
        if biases is not None:
            bi = biases.values[biases.chr_off[ci] + i]
            bj = biases.values[biases.chr_off[ci] + j]
            lam = lam * np.where(bi > 0, bi, 1.0) * np.where(bj > 0, bj, 1.0)
        c = 1 + rng.poisson(lam)
        m1s.append(i * res + res // 2)
        m2s.append(j * res + res // 2)
        cs.append(c)
        chs.append(np.full(n, ci | (ci << 16), dtype=np.uint32))
    m1 = np.concatenate(m1s)
    m2 = np.concatenate(m2s)
    cnt = np.concatenate(cs)
    ch = np.concatenate(chs)
    n_inter = int(n_pairs * inter_fraction)
    if n_inter:
        c1 = rng.integers(0, len(names), n_inter)
        c2 = (c1 + 1 + rng.integers(0, len(names) - 1, n_inter)) % len(names)
        a, b = np.minimum(c1, c2), np.maximum(c1, c2)
        i1 = np.floor(rng.random(n_inter) * nb[a]).astype(np.int64)
        i2 = np.floor(rng.random(n_inter) * nb[b]).astype(np.int64)
        m1 = np.concatenate([m1, i1 * res + res // 2])
        m2 = np.concatenate([m2, i2 * res + res // 2])
        cnt = np.concatenate([cnt, 1 + rng.poisson(0.3, n_inter)])
        ch = np.concatenate([ch, (a | (b << 16)).astype(np.uint32)])
        perm = rng.permutation(len(m1))
        m1, m2, cnt, ch = m1[perm], m2[perm], cnt[perm], ch[perm]
    contacts = Contacts(m1.astype(np.int32), m2.astype(np.int32), cnt.astype(np.int32), ch.astype(np.uint32), list(names))
    return contacts, frags, biases, raw


def lpt_shards(weights, nshards):
    """Largest-processing-time-first assignment of chromosomes to GPUs.  Returns list of index lists."""
    order = sorted(range(len(weights)), key=lambda i: -weights[i])
    loads = [0] * nshards
    out = [[] for _ in range(nshards)]
    for i in order:
        r = min(range(nshards), key=lambda k: loads[k])
        out[r].append(i)
        loads[r] += weights[i]
    return [sorted(o) for o in out]


def make_intra_device(n_pairs, res, seed, device, mean_count=3.0, with_bias=True, only=None, chunk=1 << 26, order="file",
                      chroms=None, signal_frac=0.0):
    """torch generator on the GPU for bench-sized inputs (300 M pairs in seconds).  Same law as make_intra; every
    chromosome has its own seeded stream, so a rank that generates only the chromosomes in `only` (indices) gets exactly
    the lines the single-GPU run has for them.  order = "file": the lines of a chromosome are sorted by (mid1, mid2), the
    order fixed-size-bin contact files come in (e.g. the reference's fithic/tests/data/contactCounts/*_w40000_chr1.gz:
    all partners of one locus on consecutive lines); "random": the order the pairs were drawn in.  Returns ((mid1, mid2,
    cnt, chrs) int32 device tensors, Fragments, Biases, per-chromosome pair counts of the WHOLE data set).
    chroms: restrict the genome (e.g. ["chr1"]).  signal_frac: this share of the lines (drawn per line, independent of its
    distance) gets 4 + Poisson(6) extra reads -- planted interactions, so that a comparable share of the lines ends with
    q < 1 like on real maps (the reference's bundled data: 3 ... 48 % of the lines), instead of none under the null model."""
    import torch
    names, sizes = genome(chroms)
    nb = n_bins(sizes, res)
    frags = fragments_for(names, sizes, res)
    rng = np.random.Generator(np.random.PCG64(seed))
    biases, _ = (make_biases(names, sizes, res, rng) if with_bias else (None, None))
    w = sizes / sizes.sum()
    per = np.floor(w * n_pairs).astype(np.int64)
    per[0] += n_pairs - per.sum()
    kmax = int(nb.max())
    ks = np.arange(kmax)
    pk = 1.0 / (ks + 1.0)
    pk /= pk.sum()
    lam0 = (mean_count - 1.0) / float((pk * (ks + 1.0) ** -1.08).sum())
    which = list(range(len(names))) if only is None else list(only)
    n_local = int(per[which].sum()) if len(which) else 0
    mid1 = torch.empty(n_local, dtype=torch.int32, device=device)
    mid2 = torch.empty(n_local, dtype=torch.int32, device=device)
    cnt = torch.empty(n_local, dtype=torch.int32, device=device)
    chrs = torch.empty(n_local, dtype=torch.int32, device=device)
    bvals = torch.from_numpy(biases.values).to(device) if biases is not None else None
    pos = 0
    for ci in which:
        n = int(per[ci])
        nbc = int(nb[ci])
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000 + ci)
        done = 0
        first = pos
        while done < n:
            m = min(chunk, n - done)
            u = torch.rand(m, generator=g, device=device, dtype=torch.float64)
            k = torch.clamp(torch.floor(torch.pow(torch.tensor(nbc + 1.0, device=device, dtype=torch.float64), u)).long() - 1,
                            max=nbc - 1)
            i = torch.floor(torch.rand(m, generator=g, device=device, dtype=torch.float64) * (nbc - k)).long()
            j = i + k
            lam = lam0 * torch.pow(k.double() + 1.0, -1.08)
            if bvals is not None:
                off = int(biases.chr_off[ci])
                bi = bvals[off + i]
                bj = bvals[off + j]
                lam = lam * torch.where(bi > 0, bi, torch.ones_like(bi)) * torch.where(bj > 0, bj, torch.ones_like(bj))
            c = 1 + torch.poisson(lam, generator=g).long()
            if signal_frac > 0.0:
                hit = torch.rand(m, generator=g, device=device) < signal_frac
                extra = 4 + torch.poisson(torch.full((m,), 6.0, device=device), generator=g).long()
                c = torch.where(hit, c + extra, c)
            s = slice(pos, pos + m)
            mid1[s] = (i * res + res // 2).int()
            mid2[s] = (j * res + res // 2).int()
            cnt[s] = c.int()
            chrs[s] = ci | (ci << 16)
            pos += m
            done += m
        if order == "file" and pos > first:
            key = mid1[first:pos].long() * (nbc * res + res) + mid2[first:pos].long()
            perm = torch.argsort(key)
            del key
            for t in (mid1, mid2, cnt):
                t[first:pos] = t[first:pos][perm]
            del perm
    return (mid1, mid2, cnt, chrs), frags, biases, per


def make_inter_device(n_pairs, res, seed, device, intra_fraction=0.1, chunk=1 << 24, only_chunks=None):
    """BASELINE config 5 on the GPU: whole-genome interOnly input -- chromosome pairs drawn with probability proportional to
    the product of their lengths, loci uniform, count = 1 + Poisson(0.3), no bias; `intra_fraction` of the lines are intra
    pairs (under -x interOnly the reference scores those against the inter prior as well, fithic/fithic.py:1098-1108).
    The file is a sequence of chunks of `chunk` lines, each with its own seeded stream: a rank that generates only the
    chunks in `only_chunks` (a range) gets exactly the lines the single-GPU run has there.
    Returns ((mid1, mid2, cnt, chrs) int32 device tensors, Fragments, first file line of this rank's lines)."""
    import torch
    names, sizes = genome(None)
    nb = n_bins(sizes, res)
    frags = fragments_for(names, sizes, res)
    w = torch.from_numpy(sizes / sizes.sum()).to(device)
    nbt = torch.from_numpy(nb).to(device)
    nchunks = (n_pairs + chunk - 1) // chunk
    which = range(nchunks) if only_chunks is None else only_chunks
    lo = min(which.start * chunk, n_pairs) if len(which) else 0
    hi = min(which.stop * chunk, n_pairs) if len(which) else 0
    n_local = hi - lo
    mid1 = torch.empty(n_local, dtype=torch.int32, device=device)
    mid2 = torch.empty(n_local, dtype=torch.int32, device=device)
    cnt = torch.empty(n_local, dtype=torch.int32, device=device)
    chrs = torch.empty(n_local, dtype=torch.int32, device=device)
    done = 0
    for ck in which:
        m = min(chunk, n_pairs - ck * chunk)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 100003 + ck)
        c1 = torch.multinomial(w, m, replacement=True, generator=g)
        c2 = torch.multinomial(w, m, replacement=True, generator=g)
        same = torch.rand(m, generator=g, device=device) < intra_fraction
        c2 = torch.where(same, c1, torch.where(c2 == c1, (c1 + 1) % len(names), c2))
        i = torch.floor(torch.rand(m, generator=g, device=device, dtype=torch.float64) * nbt[c1]).long()
        j = torch.floor(torch.rand(m, generator=g, device=device, dtype=torch.float64) * nbt[c2]).long()
        c = 1 + torch.poisson(torch.full((m,), 0.3, device=device), generator=g).long()
        s = slice(done, done + m)
        mid1[s] = (i * res + res // 2).int()
        mid2[s] = (j * res + res // 2).int()
        cnt[s] = c.int()
        chrs[s] = (c1 | (c2 << 16)).int()
        done += m
    return (mid1, mid2, cnt, chrs), frags, lo


def write_inputs(outdir, contacts, frags, res, raw_bias=None, biases=None, prefix="synth"):
    """Write gz TSV files in the reference's input formats (for CLI-level tests and the CPU baseline)."""
    import gzip
    import os
    os.makedirs(outdir, exist_ok=True)
    names = contacts.chroms
    cpath = os.path.join(outdir, prefix + ".contacts.gz")
    fpath = os.path.join(outdir, prefix + ".fragments.gz")
    c1 = contacts.chrs & 0xffff
    c2 = contacts.chrs >> 16
    with gzip.open(cpath, "wt", compresslevel=1) as f:
        f.write("".join("%s\t%d\t%s\t%d\t%d\n" % (names[a], m1, names[b], m2, c) for a, m1, b, m2, c in
                        zip(c1.tolist(), contacts.mid1.tolist(), c2.tolist(), contacts.mid2.tolist(),
                            contacts.cnt.tolist())))
    with gzip.open(fpath, "wt", compresslevel=1) as f:
        for ci, name in enumerate(frags.chroms):
            if getattr(frags, "mids", None) is not None:  # restriction fragments: the mid points as they are
                f.write("".join("%s\t0\t%d\t1\t1\n" % (name, m) for m in np.asarray(frags.mids[ci]).tolist()))
                continue
            n = int(frags.n_mappable[ci])
            f.write("".join("%s\t0\t%d\t1\t1\n" % (name, k * res + res // 2) for k in range(n)))
    bpath = None
    if raw_bias is None and biases is not None and getattr(biases, "sparse", False):
        raw_bias = biases.values  # a sparse table lists only what the bias file held (out-of-bounds values already -1)
    if raw_bias is not None:
        bpath = os.path.join(outdir, prefix + ".bias.gz")
        with gzip.open(bpath, "wt", compresslevel=1) as f:
            for ci, name in enumerate(frags.chroms):
                lo, hi = int(biases.chr_off[ci]), int(biases.chr_off[ci + 1])
                f.write("".join("%s\t%d\t%r\n" % (name, int(m), float(v)) for m, v in
                                zip(biases.mids[lo:hi].tolist(), raw_bias[lo:hi].tolist())))
    return cpath, fpath, bpath
