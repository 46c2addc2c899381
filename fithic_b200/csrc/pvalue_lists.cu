// K3 as a work-list pipeline -- the same arithmetic as pvalue.cu, organised so that every phase runs with full warps.
//
// Replaces the per-line loop of fit_Spline (reference fithic/fithic.py:1017-1123) and scipy.special.bdtrc (:1070, :1101).
// K3 is bound by instruction issue, not by HBM (ncu on the tile-phased kernel: issue slots 53 %, DRAM 8 %): what costs
// time is lanes that idle while their neighbours run another evaluation or a longer continued fraction.  The tile-phased
// kernel fights that inside a 2048-contact tile, but a tile only holds ~2 iterative items per lane, so its work list
// drains almost as soon as it starts (13 of 32 lanes active on average in the iteration loop).  Here the work lists live
// in HBM instead (16 B per item: 8 % of the DRAM bandwidth was in use), which makes them as long as the input:
//
//   pval_front2_kernel   every contact: classify, bias gather, prior, ExpCC; results known at once and the count == 1
//                        closed form (65 % of a sparse map) are finished here; contacts that need a continued fraction or
//                        a tail sum are appended to two lists (one shuffle scan + one global atomic per warp and 2048
//                        contacts; pval_front_kernel is the first version: CTA-wide scan between two barriers)
//   pval_iterate_kernel  persistent warps pull 128-item chunks from a list; a lane whose item has converged stores
//                        numerator/denominator and takes the next item at once, so warps stay full until the list ends;
//                        tail sums of up to 32 terms are left to the finish kernel
//   pval_finish_kernel   one thread per item, uniform: short tail sums in place, prefactor in log space (2 log + 1 exp, no
//                        division besides P/Q), p scattered to its line, outlier flag
#define FHC_PROFILE_STREAM st
#include <stdlib.h>

#include "pvalue_common.cuh"

namespace fhc {

constexpr int kFrontThreads = 256;
constexpr int kFrontTile = 2048;  // contacts per CTA iteration (two groups of four per thread)
constexpr int kIterThreads = 256;
constexpr int kIterChunk = 128;   // items a warp claims with one global atomic and stages in shared memory
constexpr int kFinishThreads = 256;

struct __align__(16) WorkItem {
    double x;          // prior
    unsigned int idx;  // line (relative to this call)
    int cnt;           // count; bit 31: scored against N_inter
};

struct ListsWs {
    unsigned long long *ctr;  // [0] list lengths: continued fractions (low word) | tail sums (high word); [2]/[3] the cursors of
                              // the iterate kernel
    WorkItem *items;          // capacity n: continued fractions from the front, tail sums from the back
    double2 *pq;              // numerator / denominator of the item at the same position
    double2 *aux_intra, *aux_inter;  // per count: lbeta + log(count), lbeta + log(N - count + 1)
    long long cap;
};

// priors of 2^-17 and above are rare (dense short-range bins): one out-of-line copy of the general closed form
__device__ __noinline__ double bdtrc_k0_slow(int N, double prior) { return bdtrc_k0(N, prior); }

__device__ __forceinline__ double bdtrc_k0_fast(int N, double prior) {
    return k0_series_ok(prior) ? bdtrc_k0_series((double)N, prior) : bdtrc_k0_slow(N, prior);
}

// ---- the per-line branch order of fit_Spline (fithic/fithic.py:1057-1115) and the exits of scipy.special.bdtrc /
// cephes incbet before any real work, written as selects instead of early returns: one straight instruction stream for
// all 32 lanes (the early-return form of pvalue_common.cuh compiles to ~25 branches per contact).
struct FrontConst {
    unsigned int Llo, Uhi;  // in-range window clamped to what a 32-bit distance can reach
    bool nothing_in_range;  // L beyond 2^32 - 1
    double dN_intra, dN_inter, dNp1_intra, dNp1_inter;
    // n / res for n < 2^31 by one 32-bit multiply-high: m = ceil(2^(31 + l) / res), l = ceil(log2 res), q = hi32(n m) >> (l - 1)
    // (e = m res - 2^(31 + l) < res <= 2^l and n < 2^31 give n e < 2^(31 + l): exact); res == 1 is flagged
    unsigned int div_m, div_sh;
    unsigned int D32;   // table length clamped to 2^32 - 1; 0 when there is no table (no slot is ever "inside" then)
    unsigned int half;  // res / 2: where the mid point of a locus of the regular grid sits inside its bin
};

__device__ __forceinline__ unsigned int fastdiv31(unsigned int n, const PvalParams &P, const FrontConst &F) {
    return P.res.d == 1 ? n : (__umulhi(n, F.div_m) >> F.div_sh);
}

constexpr int kChrSmem = 1024;  // chromosome slot ranges kept in shared memory

// REGULAR: the slots hold the loci of the regular grid (P.bias_mid == nullptr): the mid point is checked arithmetically.
// chr_rng[c] = [first slot, end) of chromosome c as 32-bit values (the caller falls back to bias_lookup of
// pvalue_common.cuh when there are more than kChrSmem chromosomes or 2^31 slots).
template <bool REGULAR>
__device__ __forceinline__ double bias_lookup_sel(const PvalParams &P, const FrontConst &F, const int2 *chr_rng,
                                                  unsigned int chr, int mid) {
    bool ok = chr < (unsigned int)P.nchr && mid >= 0;
    const int2 rng = chr_rng[ok ? chr : 0u];
    const unsigned int k = fastdiv31((unsigned int)mid, P, F);  // garbage for mid < 0, masked by ok
    int s = rng.x + (int)k;
    ok = ok && s < rng.y && s >= rng.x;
    s = ok ? s : 0;
    if (REGULAR)
        ok = ok && ((unsigned int)mid - k * P.res.d == (P.res.d >> 1));
    else
        ok = ok && __ldg(P.bias_mid + s) == mid;
    const double b = __ldg(P.bias + s);
    return ok ? b : -1.0;
}

// the same when the chromosome -- and with it the slot range -- is known for the whole tile
template <bool REGULAR>
__device__ __forceinline__ double bias_lookup_rng(const PvalParams &P, const FrontConst &F, int2 rng, bool chr_ok, int mid) {
    bool ok = chr_ok && mid >= 0;
    const unsigned int k = fastdiv31((unsigned int)mid, P, F);  // garbage for mid < 0, masked by ok
    int s = rng.x + (int)k;
    ok = ok && s < rng.y && s >= rng.x;
    s = ok ? s : 0;
    if (REGULAR)
        ok = ok && ((unsigned int)mid - k * P.res.d == (P.res.d >> 1));
    else
        ok = ok && __ldg(P.bias_mid + s) == mid;
    const double b = __ldg(P.bias + s);
    return ok ? b : -1.0;
}

// the lookup for sparse bias tables (restriction fragments: binary search), more than kChrSmem chromosomes or 2^31 slots:
// out of line, so that its four inlined copies per group do not sit in the middle of the regular grid's instruction stream
// (the fields it needs go by value: a reference to the kernel's parameter block would force a copy of it onto the stack)
__device__ __noinline__ double bias_lookup_general_fn(const long long *chr_off, const int *bias_mid, const double *bias,
                                                      int bias_sparse, int nchr, FastDiv res, unsigned int chr, int mid) {
    PvalParams Q;
    Q.chr_off = chr_off;
    Q.bias_mid = bias_mid;
    Q.bias = bias;
    Q.bias_sparse = bias_sparse;
    Q.nchr = nchr;
    Q.res = res;
    return bias_lookup(Q, chr, mid);
}
__device__ __forceinline__ double bias_lookup_general(const PvalParams &P, unsigned int chr, int mid) {
    return bias_lookup_general_fn(P.chr_off, P.bias_mid, P.bias, P.bias_sparse, P.nchr, P.res, chr, mid);
}

// phase A of a contact: the three gathers (two bias values, the distance table), issued for all four contacts of a group
// before anything consumes them so that their L2 latencies overlap
// INTRA: every contact of the tile lies on one chromosome (known from the chromosome runs): `ch` is that chromosome's
// id pair, `rng` / `chr_ok` its slot range in the bias table, and nothing per contact depends on chromosome ids.
template <bool HAS_BIAS, bool REGULAR, bool INTRA>
__device__ __forceinline__ void front_gather(const PvalParams &P, const FrontConst &F, const int2 *chr_rng, bool rng32, int2 rng,
                                             bool chr_ok, int m1, int m2, unsigned int ch, double &b1, double &b2, double &tabv,
                                             unsigned int &d) {
    const unsigned int c1 = ch & 0xffffu, c2 = ch >> 16;
    d = m1 > m2 ? (unsigned int)m1 - (unsigned int)m2 : (unsigned int)m2 - (unsigned int)m1;
    b1 = 1.0;
    b2 = 1.0;
    if (HAS_BIAS) {
        if (rng32) {  // uniform
            if (INTRA) {
                b1 = bias_lookup_rng<REGULAR>(P, F, rng, chr_ok, m1);
                b2 = bias_lookup_rng<REGULAR>(P, F, rng, chr_ok, m2);
            } else {
                b1 = bias_lookup_sel<REGULAR>(P, F, chr_rng, c1, m1);
                b2 = bias_lookup_sel<REGULAR>(P, F, chr_rng, c2, m2);
            }
        } else {
            b1 = bias_lookup_general(P, c1, m1);
            b2 = bias_lookup_general(P, c2, m2);
        }
    }
    const unsigned int slot = d < 0x80000000u ? fastdiv31(d, P, F) : fastdiv(d, P.res);
    const bool slot_ok = (INTRA || c1 == c2) && slot < F.D32;
    tabv = NAN;  // beyond the table (or an inter line, which never uses it)
    if (P.lut != nullptr) {  // uniform branch; the index is always valid
        const double t = __ldg(P.lut + (slot_ok ? slot : 0u));
        tabv = slot_ok ? t : NAN;
    }
}

// phase B, first half: what a contact is before any table value is known (the branch order of fithic/fithic.py:1057-1115)
// INTRA: an intra tile of a run that is not interOnly (the caller checks the mode): the intra branch with constants folded
struct FrontFlags {
    bool scored, intra_path, b_ok;
};
template <bool INTRA>
__device__ __forceinline__ FrontFlags front_flags(const PvalParams &P, const FrontConst &F, unsigned int d, unsigned int ch,
                                                  bool in_file, double b1, double b2) {
    const bool inter = INTRA ? false : (ch & 0xffffu) != (ch >> 16);
    FrontFlags f;
    f.intra_path = INTRA ? true : (!inter && P.mode != FHC_MODE_INTER_ONLY);
    // :1057-1063.  Two compares chained through the predicate (written as b1 < 0 || b2 < 0 the compiler builds
    // min(b1, b2) with its NaN rules out of nine instructions)
    unsigned int neg;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, 0d0000000000000000;\n\tsetp.lt.or.f64 p, %2, 0d0000000000000000, p;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(neg)
        : "d"(b1), "d"(b2));
    const bool discarded = neg != 0u && !inter;
    const bool in_range = d >= F.Llo && d <= F.Uhi && !F.nothing_in_range;                      // :1065 / :1081-1096
    f.scored = in_file && !discarded && (f.intra_path ? in_range : P.mode != FHC_MODE_INTRA_ONLY);
    f.b_ok = b1 >= P.tL && b1 <= P.tU && b2 >= P.tL && b2 <= P.tU;
    return f;
}

// value of a scored line that needs neither the closed form nor an iteration (lowest priority first)
__device__ __noinline__ double front_done_value(int c, unsigned int k, unsigned int N, double prior) {
    double v = prior >= 1.0 ? 1.0 : 0.0;               // incbet: xx == 1 -> 1, xx == 0 -> 0
    if (k == N) v = 0.0;                               // bdtrc: k == n
    if (k > N) v = NAN;                                // bdtrc: n < k
    if (c <= 0) v = 1.0;                               // bdtrc: k < 0
    if (!(prior >= 0.0 && prior <= 1.0)) v = NAN;      // NaN or outside [0, 1]
    return v;
}

// phase B, second half: prior, ExpCC and the exits of bdtrc / incbet before any real work.  b12 = rn(b1 b2).
__device__ __forceinline__ PvalClass front_score(const PvalParams &P, const FrontConst &F, const FrontFlags f, int c, double b12,
                                                 double tabv, double &p, double &e, double &prior, bool &use_inter) {
    const bool intra_path = f.intra_path, scored = f.scored;
    use_inter = !intra_path;
    const double prior0 = intra_path ? tabv : P.interChrProb;  // tabv is NaN beyond the table
    prior = __dmul_rn(prior0, b12);
    const double dN = intra_path ? F.dN_intra : F.dN_inter;
    const unsigned int N = (unsigned int)(intra_path ? P.N_intra : P.N_inter);  // 0 <= N < 2^31
    // ExpCC = N prior where the line is scored and both biases lie in [tL, tU], else +0.0: as a bit mask (a conditional
    // on doubles compiles to one pair of selects per term of the condition)
    // (the empty asm keeps the compiler from folding the mask back into that form)
    unsigned int eok = (scored && f.b_ok) ? 0xffffffffu : 0u;
    asm("" : "+r"(eok));
    const double en = __dmul_rn(dN, prior);
    e = __hiloint2double(__double2hiint(en) & (int)eok, __double2loint(en) & (int)eok);
    // bdtrc(k = c - 1, N, prior) and incbet(c, N - c + 1, prior) up to the first real work (cephes bdtr.h / incbet.h;
    // bdtrc_classify in cephes_dev.cuh is the same ladder with early returns).  k = c - 1 is compared as an unsigned 32-bit
    // value (c <= 0 wraps to a huge k).  The two classes with real work are decided directly:
    //   count == 1: the closed form, unless N == 0 (k == N) or the prior is NaN / outside [0, 1]
    //   count >= 2: an iteration, when k < N and 0 < prior < 1
    // and everything else that is scored -- rare: counts beyond N, priors of exactly 0 or 1, NaN priors beyond the table --
    // gets its value from the ladder in front_done_value.
    const unsigned int k = (unsigned int)c - 1u;
    const bool k0 = scored & (c == 1) & (N != 0u) & (prior >= 0.0) & (prior <= 1.0);
    const bool iter = scored & (c >= 2) & (k < N) & (prior > 0.0) & (prior < 1.0);
    PvalClass cls = k0 ? kClsK0 : kClsDone;
    if (iter) {
        const double dNp1 = intra_path ? F.dNp1_intra : F.dNp1_inter;
        cls = __dmul_rn(prior, dNp1) > (double)c ? kClsTail : kClsCf;  // x > a / (a + b), a + b = N + 1
    }
    p = 1.0;  // not scored: p = 1
    if (scored && !k0 && !iter) p = front_done_value(c, k, N, prior);
    return cls;
}

template <bool INTRA>
__device__ __forceinline__ PvalClass front_prepare(const PvalParams &P, const FrontConst &F, unsigned int d, int c,
                                                   unsigned int ch, bool in_file, double b1, double b2, double tabv,
                                                   double &p, double &e, double &prior, bool &use_inter) {
    const FrontFlags f = front_flags<INTRA>(P, F, d, ch, in_file, b1, b2);
    return front_score(P, F, f, c, __dmul_rn(b1, b2), tabv, p, e, prior, use_inter);
}

// What the pre-pass leaves per contact (fhc_pvalues_prepass): everything of the per-line branch order that does not need
// the spline table, computed while the host bins and fits.  code: kPreNotScored, or b_ok << 31 | scored-against-the-inter-
// prior << 30 | distance slot (intra path; kPreSlotMask = beyond any table); b12 = rn(bias1 bias2).
constexpr unsigned int kPreNotScored = 0xffffffffu, kPreBok = 0x80000000u, kPreInter = 0x40000000u, kPreSlotMask = 0x3fffffffu;

// ---- front ------------------------------------------------------------------------------------------------------------
struct FrontSmem {
    long long run_start[FHC_MAX_CHR_RUNS + 1];  // chromosome ids as runs (P.chrs == nullptr): run r = lines [rs[r], rs[r+1])
    unsigned int run_val[FHC_MAX_CHR_RUNS];
    int2 chr_rng[kChrSmem];  // [first slot, end) of each chromosome in the dense bias table
    double x[kFrontTile];
    int cnt[kFrontTile];  // count | inter << 31
    unsigned int warp_tot[kFrontThreads / 32];
    unsigned long long base_cf, base_tail;
};

// chromosome id pair of one line from the run table (binary search; only tiles that straddle a run boundary come here)
__device__ __forceinline__ unsigned int run_lookup(const FrontSmem &S, int nruns, long long line) {
    int lo = 0, hi = nruns - 1;  // last run that starts at or before `line`
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (S.run_start[mid] <= line)
            lo = mid;
        else
            hi = mid - 1;
    }
    return S.run_val[lo];
}

// One tile of kFrontTile contacts: loads, gathers, classification, closed forms, stores of p and ExpCC; the contacts that
// need iterating are parked in S.x / S.cnt.  Returns the 2-bit classes of this thread's 8 contacts.
// kG contacts per thread and load (4: 128-bit loads and stores, 80 registers, 3 CTAs per SM; 2: 64-bit loads, 128-bit
// stores of two doubles, fits 64 registers, 4 CTAs per SM); a thread handles 8 contacts of a tile either way.
// INTRA: the whole (full) tile lies in one intra chromosome run: no chromosome ids per contact at all.
// One tile of the kernel that follows a pre-pass (P.pre_code / P.pre_b12): the classification and the bias product come
// from there, what is left is the table lookup, the prior and everything behind it.
template <int kG>
__device__ __forceinline__ unsigned int front_tile_pre(const PvalParams &P, const FrontConst &F, FrontSmem &S, long long base,
                                                       bool full, unsigned int &flagged) {
    const int tid = threadIdx.x;
    const int *cs = reinterpret_cast<const int *>(P.cnt);
    unsigned int codes = 0;
#pragma unroll
    for (int h = 0; h < 8 / kG; ++h) {
        const int l0 = (h * kFrontThreads + tid) * kG;
        int cc[kG];
        unsigned int pc[kG];
        double b12[kG];
        if (full) {
            static_assert(kG == 2, "the pre-pass variant loads two contacts per thread and step");
            const long long g = (base + l0) >> 1;
            const int2 ac = ldg_stream2(reinterpret_cast<const int2 *>(P.cnt) + g);
            const int2 ap = ldg_stream2(reinterpret_cast<const int2 *>(P.pre_code) + g);
            const double2 ab = __ldcs(reinterpret_cast<const double2 *>(P.pre_b12) + g);
            cc[0] = ac.x; cc[1] = ac.y;
            pc[0] = (unsigned int)ap.x; pc[1] = (unsigned int)ap.y;
            b12[0] = ab.x; b12[1] = ab.y;
        } else {
#pragma unroll
            for (int k = 0; k < kG; ++k) {
                const long long i = base + l0 + k;
                const bool ok = i < P.n;
                cc[k] = ok ? cs[i] : 0;
                pc[k] = ok ? P.pre_code[i] : kPreNotScored;
                b12[k] = ok ? P.pre_b12[i] : 1.0;
            }
        }
        double gtv[kG];
#pragma unroll
        for (int k = 0; k < kG; ++k) {  // the one gather that is left
            const unsigned int slot = pc[k] & kPreSlotMask;
            const bool want = pc[k] != kPreNotScored && !(pc[k] & kPreInter) && slot < F.D32 && P.lut != nullptr;
            const double t = P.lut != nullptr ? __ldg(P.lut + (want ? slot : 0u)) : 0.0;
            gtv[k] = want ? t : NAN;
        }
        double e[kG], pv[kG];
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            const int li = l0 + k;
            FrontFlags f;
            f.scored = pc[k] != kPreNotScored;
            f.intra_path = !(pc[k] & kPreInter);
            f.b_ok = (pc[k] & kPreBok) != 0;
            double prior;
            bool use_inter;
            const PvalClass cls = front_score(P, F, f, cc[k], b12[k], gtv[k], pv[k], e[k], prior, use_inter);
            if (cls == kClsK0) {
                pv[k] = bdtrc_k0_fast(use_inter ? P.N_inter : P.N_intra, prior);
            } else if (cls != kClsDone) {
                S.x[li] = prior;
                S.cnt[li] = cc[k] | (use_inter ? (int)0x80000000u : 0);
                pv[k] = 0.0;  // overwritten by pval_finish_kernel
            }
            codes |= (unsigned int)cls << (2 * (h * kG + k));
        }
        if (full) {
            double2 *ee = reinterpret_cast<double2 *>(P.expcc + base + l0);
            double2 *pp = reinterpret_cast<double2 *>(P.p + base + l0);
            __stcs(ee, make_double2(e[0], e[1]));
            pp[0] = make_double2(pv[0], pv[1]);
        } else {
#pragma unroll
            for (int k = 0; k < kG; ++k)
                if (base + l0 + k < P.n) {
                    P.expcc[base + l0 + k] = e[k];
                    P.p[base + l0 + k] = pv[k];
                }
        }
        if (P.outl != nullptr) {
#pragma unroll
            for (int k = 0; k < kG; ++k) {
                const unsigned int c = (codes >> (2 * (h * kG + k))) & 3u;
                if ((c == kClsDone || c == kClsK0) && base + l0 + k < P.n) outlier_mark(P, base + l0 + k, pv[k], flagged);
            }
        }
    }
    return codes;
}

template <bool HAS_BIAS, bool REGULAR, int kG, bool INTRA>
__device__ __forceinline__ unsigned int front_tile(const PvalParams &P, const FrontConst &F, FrontSmem &S, bool rng32,
                                                   long long base, bool full, bool tile_one_run, unsigned int tile_ch,
                                                   unsigned int &flagged) {
    const int tid = threadIdx.x;
    const int *m1s = reinterpret_cast<const int *>(P.mid1), *m2s = reinterpret_cast<const int *>(P.mid2);
    const int *cs = reinterpret_cast<const int *>(P.cnt);
    const unsigned int *hs = reinterpret_cast<const unsigned int *>(P.chrs);
    int2 rng = make_int2(0, 0);
    bool chr_ok = false;
    if (INTRA && HAS_BIAS && rng32) {
        const unsigned int c = tile_ch & 0xffffu;
        chr_ok = c < (unsigned int)P.nchr;
        rng = S.chr_rng[chr_ok ? c : 0u];
    }
    unsigned int codes = 0;  // 2 bits per contact of this thread: PvalClass
#pragma unroll
    for (int h = 0; h < 8 / kG; ++h) {
        const int l0 = (h * kFrontThreads + tid) * kG;
        int m1[kG], m2[kG], cc[kG];
        unsigned int ch[kG];
        if (full) {
            if (kG == 4) {
                const long long g = (base + l0) >> 2;
                const int4 a1 = ldg_stream(P.mid1 + g), a2 = ldg_stream(P.mid2 + g), ac = ldg_stream(P.cnt + g);
                m1[0] = a1.x; m1[1] = a1.y; m1[kG - 2] = a1.z; m1[kG - 1] = a1.w;
                m2[0] = a2.x; m2[1] = a2.y; m2[kG - 2] = a2.z; m2[kG - 1] = a2.w;
                cc[0] = ac.x; cc[1] = ac.y; cc[kG - 2] = ac.z; cc[kG - 1] = ac.w;
                if (!INTRA && P.chrs != nullptr) {
                    const int4 ah = ldg_stream(P.chrs + g);
                    ch[0] = (unsigned int)ah.x; ch[1] = (unsigned int)ah.y;
                    ch[kG - 2] = (unsigned int)ah.z; ch[kG - 1] = (unsigned int)ah.w;
                }
            } else {
                const long long g = (base + l0) >> 1;
                const int2 a1 = ldg_stream2(reinterpret_cast<const int2 *>(P.mid1) + g);
                const int2 a2 = ldg_stream2(reinterpret_cast<const int2 *>(P.mid2) + g);
                const int2 ac = ldg_stream2(reinterpret_cast<const int2 *>(P.cnt) + g);
                m1[0] = a1.x; m1[1] = a1.y;
                m2[0] = a2.x; m2[1] = a2.y;
                cc[0] = ac.x; cc[1] = ac.y;
                if (!INTRA && P.chrs != nullptr) {
                    const int2 ah = ldg_stream2(reinterpret_cast<const int2 *>(P.chrs) + g);
                    ch[0] = (unsigned int)ah.x; ch[1] = (unsigned int)ah.y;
                }
            }
            if (INTRA || P.chrs == nullptr) {
#pragma unroll
                for (int k = 0; k < kG; ++k)
                    ch[k] = (INTRA || tile_one_run) ? tile_ch : run_lookup(S, P.nruns, P.line_base + base + l0 + k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kG; ++k) {
                const long long i = base + l0 + k;
                const bool ok = i < P.n;
                m1[k] = ok ? m1s[i] : 0;
                m2[k] = ok ? m2s[i] : 0;
                cc[k] = ok ? cs[i] : 0;
                ch[k] = 0x00010000u;  // padding: an inter line
                if (ok) ch[k] = P.chrs != nullptr ? hs[i] : (tile_one_run ? tile_ch : run_lookup(S, P.nruns, P.line_base + i));
            }
        }
        double e[kG], pv[kG], gb1[kG], gb2[kG], gtv[kG];
        unsigned int dd[kG];
#pragma unroll
        for (int k = 0; k < kG; ++k)
            front_gather<HAS_BIAS, REGULAR, INTRA>(P, F, S.chr_rng, rng32, rng, chr_ok, m1[k], m2[k], ch[k], gb1[k], gb2[k],
                                                   gtv[k], dd[k]);
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            const int li = l0 + k;
            double prior;
            bool use_inter;
            const bool in_file = full || base + li < P.n;
            const PvalClass cls = front_prepare<INTRA>(P, F, dd[k], cc[k], ch[k], in_file, gb1[k], gb2[k], gtv[k], pv[k], e[k],
                                                       prior, use_inter);
            if (cls == kClsK0) {
                pv[k] = bdtrc_k0_fast(use_inter ? P.N_inter : P.N_intra, prior);
            } else if (cls != kClsDone) {
                S.x[li] = prior;
                S.cnt[li] = cc[k] | (use_inter ? (int)0x80000000u : 0);
                pv[k] = 0.0;  // overwritten by pval_finish_kernel
            }
            codes |= (unsigned int)cls << (2 * (h * kG + k));
        }
        if (full) {
            double2 *ee = reinterpret_cast<double2 *>(P.expcc + base + l0);
            double2 *pp = reinterpret_cast<double2 *>(P.p + base + l0);
#pragma unroll
            for (int k = 0; k < kG; k += 2) {
                __stcs(ee + k / 2, make_double2(e[k], e[k + 1]));
                pp[k / 2] = make_double2(pv[k], pv[k + 1]);  // default caching: the finish kernel writes into these lines soon
            }
        } else {
#pragma unroll
            for (int k = 0; k < kG; ++k)
                if (base + l0 + k < P.n) {
                    P.expcc[base + l0 + k] = e[k];
                    P.p[base + l0 + k] = pv[k];
                }
        }
        if (P.outl != nullptr) {
#pragma unroll
            for (int k = 0; k < kG; ++k) {
                const unsigned int c = (codes >> (2 * (h * kG + k))) & 3u;
                if ((c == kClsDone || c == kClsK0) && base + l0 + k < P.n) outlier_mark(P, base + l0 + k, pv[k], flagged);
            }
        }
    }
    return codes;
}

template <bool HAS_BIAS, bool REGULAR, int kMinCtas, int kG, bool PRE = false>
__global__ void __launch_bounds__(kFrontThreads, kMinCtas) pval_front_kernel(const PvalParams P, const FrontConst F,
                                                                             const ListsWs W) {
    extern __shared__ __align__(16) unsigned char front_smem[];
    FrontSmem &S = *reinterpret_cast<FrontSmem *>(front_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool rng32 = false;
    if (HAS_BIAS && !PRE) {
        rng32 = !P.bias_sparse && P.nchr <= kChrSmem && P.chr_off[P.nchr] < 0x7fffffffll;
        if (rng32)
            for (int c = tid; c < P.nchr; c += kFrontThreads) S.chr_rng[c] = make_int2((int)P.chr_off[c], (int)P.chr_off[c + 1]);
    }
    const bool runs = !PRE && P.chrs == nullptr;
    if (runs) {
        for (int r = tid; r <= P.nruns; r += kFrontThreads) S.run_start[r] = P.run_start[r];
        for (int r = tid; r < P.nruns; r += kFrontThreads) S.run_val[r] = P.run_val[r];
    }
    __syncthreads();
    const long long ntiles = (P.n + kFrontTile - 1) / kFrontTile;
    unsigned int flagged = 0;
    int run = 0;  // the tiles of this CTA only move forward, so does its position in the run table

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * kFrontTile;
        const bool full = base + kFrontTile <= P.n;
        bool one_run = false;
        unsigned int tile_ch = 0;
        if (runs) {
            const long long g0 = P.line_base + base;
            while (S.run_start[run + 1] <= g0) ++run;
            const long long g1 = P.line_base + (full ? base + kFrontTile : P.n);
            one_run = S.run_start[run + 1] >= g1;
            tile_ch = S.run_val[run];
        }
        unsigned int codes;
        if (PRE)
            codes = front_tile_pre<2>(P, F, S, base, full, flagged);
        else if (one_run && full && (tile_ch & 0xffffu) == (tile_ch >> 16) && P.mode != FHC_MODE_INTER_ONLY)
            codes = front_tile<HAS_BIAS, REGULAR, kG, true>(P, F, S, rng32, base, full, one_run, tile_ch, flagged);
        else
            codes = front_tile<HAS_BIAS, REGULAR, kG, false>(P, F, S, rng32, base, full, one_run, tile_ch, flagged);
        // positions in the two lists: CTA-wide exclusive scan of (continued fractions | tail sums << 16) per thread
        unsigned int mine = 0;
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
            const unsigned int c = (codes >> (2 * s8)) & 3u;
            mine += (c == kClsCf ? 1u : 0u) + (c == kClsTail ? 0x10000u : 0u);
        }
        unsigned int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) S.warp_tot[warp] = inc;
        __syncthreads();
        unsigned int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kFrontThreads / 32; ++w) {
            const unsigned int t = S.warp_tot[w];
            if (w < warp) pre += t;
            tot += t;
        }
        if (tid == 0) {
            const unsigned int nCf = tot & 0xffffu, nTail = tot >> 16;
            // both list lengths advance with one atomic: continued fractions in the low, tail sums in the high word
            const unsigned long long b =
                tot ? atomicAdd(W.ctr, (unsigned long long)nCf | ((unsigned long long)nTail << 32)) : 0ull;
            S.base_cf = b & 0xffffffffull;
            S.base_tail = b >> 32;
        }
        __syncthreads();
        if (mine) {
            const unsigned int ex = pre + inc - mine;
            long long oCf = (long long)S.base_cf + (ex & 0xffffu);
            long long oTail = W.cap - 1 - ((long long)S.base_tail + (ex >> 16));
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) {
                const unsigned int c = (codes >> (2 * s8)) & 3u;
                if (c == kClsCf || c == kClsTail) {
                    const int li = ((s8 / kG) * kFrontThreads + tid) * kG + (s8 % kG);
                    WorkItem it;
                    it.x = S.x[li];
                    it.idx = (unsigned int)(base + li);
                    it.cnt = S.cnt[li];
                    if (c == kClsCf)
                        W.items[oCf++] = it;
                    else
                        W.items[oTail--] = it;
                }
            }
        }
        __syncthreads();  // S.x / S.cnt / warp_tot are reused by the next tile
    }
    if (P.outl != nullptr) {
        const unsigned long long f = warp_sum((unsigned long long)flagged);
        if (lane == 0 && f) atomicAdd(P.outl_stats, f);
    }
}

static_assert(sizeof(FrontSmem) <= 48 * 1024, "the front kernel's shared memory must fit the default 48 KB");

// ---- front, second version ----------------------------------------------------------------------------------------------
// The same arithmetic per contact (front_gather / front_prepare / front_score above), two changes around it, both read off
// the per-instruction stall samples of the first version (ncu, 300 M contacts: 18 % of all samples sat on the first use of
// the streamed mid points -- once per group of two contacts, four exposed memory latencies per tile and thread --, 11 % on
// the two CTA-wide barriers around the list atomics, where 256 threads wait for one thread's round trip to L2):
//   * the streamed words of the NEXT group (and, at the end of a tile, of the CTA's next tile) are loaded into registers
//     before the current group is worked on, so their latency hides behind ~500 instructions of arithmetic;
//   * list positions are handed out per WARP: an inclusive shuffle scan of the lanes' item counts, one 64-bit atomic that
//     advances both list lengths at once (continued fractions in the low word, tail sums in the high word of ctr[0]), no
//     shared-memory exchange and no __syncthreads in the tile loop at all.  The parked items (S.x / S.cnt) are read back by
//     the thread that wrote them.  The order of the items in the lists changes; nothing downstream depends on it (each
//     item carries its line).
// The gathers of a group in two phases.  Phase 1 turns every contact of the group into where to load from: the slot of
// each bias value and of the table value, or -- where the first version selects a constant AFTER the load (locus not in the
// bias table: -1, distance beyond the table: NaN) -- the address of that constant, so the select sits BEFORE the load and
// nothing has to wait for a gathered value until the arithmetic needs it.  Phase 2 issues all loads of the group back to
// back (ncu on the form with one select behind each load: a third of the kernel's stall samples sat on those selects, the
// gathers of the second contact were only issued once the first contact's had come back).
__device__ const unsigned long long g_front_consts[2] = {0xBFF0000000000000ull, 0x7FF8000000000000ull};  // -1.0, NaN

template <bool REGULAR>
__device__ __forceinline__ const double *bias_addr_rng(const PvalParams &P, const FrontConst &F, int2 rng, bool chr_ok, int mid,
                                                       const double *minus_one) {
    // the checks of bias_lookup_rng in unsigned arithmetic: with mid >= 0 the slot rng.x + mid / res cannot wrap
    const unsigned int k = fastdiv31((unsigned int)mid, P, F);  // garbage for mid < 0, masked below
    const unsigned int s = (unsigned int)rng.x + k;
    bool ok = chr_ok & (mid >= 0) & (s < (unsigned int)rng.y);  // (no short circuits: no branches)
    if (REGULAR)
        ok = ok & ((unsigned int)mid - k * P.res.d == F.half);
    else
        ok = ok && __ldg(P.bias_mid + (ok ? s : 0u)) == mid;
    return ok ? P.bias + s : minus_one;
}

template <bool HAS_BIAS, bool REGULAR, bool INTRA>
__device__ __forceinline__ void front_addr2(const PvalParams &P, const FrontConst &F, const int2 *chr_rng, bool rng32, int2 rng,
                                            bool chr_ok, int m1, int m2, unsigned int ch, const double *&a1, const double *&a2,
                                            const double *&at, unsigned int &d) {
    const double *cst = reinterpret_cast<const double *>(g_front_consts);
    const unsigned int c1 = ch & 0xffffu, c2 = ch >> 16;
    d = m1 > m2 ? (unsigned int)m1 - (unsigned int)m2 : (unsigned int)m2 - (unsigned int)m1;
    a1 = a2 = cst;
    if (HAS_BIAS && rng32) {  // uniform
        int2 r1 = rng, r2 = rng;
        bool ok1 = chr_ok, ok2 = chr_ok;
        if (!INTRA) {
            ok1 = c1 < (unsigned int)P.nchr;
            ok2 = c2 < (unsigned int)P.nchr;
            r1 = chr_rng[ok1 ? c1 : 0u];
            r2 = chr_rng[ok2 ? c2 : 0u];
        }
        a1 = bias_addr_rng<REGULAR>(P, F, r1, ok1, m1, cst);
        a2 = bias_addr_rng<REGULAR>(P, F, r2, ok2, m2, cst);
    }
    // the launch code only takes this kernel when the table ends at or below distance 2^31, so a distance of 2^31 or more
    // (one mid point negative) lies beyond it whatever the quotient is; F.D32 is 0 when there is no table
    const unsigned int slot = fastdiv31(d, P, F);
    const bool slot_ok = (INTRA || c1 == c2) & (d < 0x80000000u) & (slot < F.D32);
    at = slot_ok ? P.lut + slot : cst + 1;  // NaN beyond the table (or an inter line, which never uses it)
}

struct StreamRegs {
    int2 a, b, c;  // mid1, mid2, count of two contacts (after a pre-pass: count, code, -)
    double2 d;     // after a pre-pass: the two bias products
};

template <bool PRE>
__device__ __forceinline__ void stream_issue(const PvalParams &P, unsigned int pair, StreamRegs &r) {  // n < 2^32
    if (PRE) {
        r.a = ldg_stream2(reinterpret_cast<const int2 *>(P.cnt) + pair);
        r.b = ldg_stream2(reinterpret_cast<const int2 *>(P.pre_code) + pair);
        r.d = ldg_stream_d2(reinterpret_cast<const double2 *>(P.pre_b12) + pair);
    } else {
        r.a = ldg_stream2(reinterpret_cast<const int2 *>(P.mid1) + pair);
        r.b = ldg_stream2(reinterpret_cast<const int2 *>(P.mid2) + pair);
        r.c = ldg_stream2(reinterpret_cast<const int2 *>(P.cnt) + pair);
    }
}

struct Front2Smem {  // (the run table is read from global memory: a few uniform, cached loads per tile)
    int2 chr_rng[kChrSmem];
    double x[kFrontTile];  // parked items: slot li belongs to the thread that handles contact li of the tile
    int cnt[kFrontTile];
};
static_assert(sizeof(Front2Smem) <= 48 * 1024, "the front kernel's shared memory must fit the default 48 KB");

__device__ __noinline__ unsigned int run_lookup2(const long long *run_start, const unsigned int *run_val, int nruns,
                                                 long long line) {
    int lo = 0, hi = nruns - 1;  // last run that starts at or before `line`
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(run_start + mid) <= line)
            lo = mid;
        else
            hi = mid - 1;
    }
    return __ldg(run_val + lo);
}

// what follows the classification of two contacts: closed form, parking, stores, outlier marks.  Returns their 4 code bits.
// i0: index of the first of the two contacts in this call's arrays (n < 2^32), l0: its slot in the tile.
__device__ __forceinline__ unsigned int front2_finish_pair(const PvalParams &P, Front2Smem &S, unsigned int i0, int l0, bool full,
                                                           const int *cc, const PvalClass *cls, const double *prior,
                                                           const bool *use_inter, double *pv, const double *e,
                                                           unsigned int &flagged) {
    const unsigned int n32 = (unsigned int)P.n;
    unsigned int codes = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int li = l0 + k;
        if (cls[k] == kClsK0) {
            pv[k] = bdtrc_k0_fast(use_inter[k] ? P.N_inter : P.N_intra, prior[k]);
        } else if (cls[k] != kClsDone) {
            S.x[li] = prior[k];
            S.cnt[li] = cc[k] | (use_inter[k] ? (int)0x80000000u : 0);
            pv[k] = 0.0;  // overwritten by pval_finish_kernel
        }
        codes |= (unsigned int)cls[k] << (2 * k);
    }
    if (full) {
        __stcs(reinterpret_cast<double2 *>(P.expcc + i0), make_double2(e[0], e[1]));
        *reinterpret_cast<double2 *>(P.p + i0) = make_double2(pv[0], pv[1]);  // default caching: see front_tile
    } else {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (i0 + k < n32) {
                P.expcc[i0 + k] = e[k];
                P.p[i0 + k] = pv[k];
            }
    }
    if (P.outl != nullptr) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if ((cls[k] == kClsDone || cls[k] == kClsK0) && (full || i0 + k < n32))
                outlier_mark(P, (long long)(i0 + k), pv[k], flagged);
    }
    return codes;
}

// One tile.  `nx` holds the streamed words of this tile's first group when the tile is full (loaded by the previous tile or
// the prologue) and leaves with those of the first group of the tile at next_base (-1: nothing to load ahead).
// pair0 = (first contact of the tile) / 2 + tid: this thread's pair of group 0; next_pair0: the same of the CTA's next tile.
template <bool HAS_BIAS, bool REGULAR, bool INTRA, bool PRE>
__device__ __forceinline__ unsigned int front2_tile(const PvalParams &P, const FrontConst &F, Front2Smem &S, bool rng32,
                                                    unsigned int pair0, bool full, bool tile_one_run, unsigned int tile_ch,
                                                    StreamRegs &nx, unsigned int next_pair0, unsigned int &flagged) {
    const int tid = threadIdx.x;
    const unsigned int n32 = (unsigned int)P.n;
    int2 rng = make_int2(0, 0);
    bool chr_ok = false;
    if (!PRE && INTRA && HAS_BIAS && rng32) {
        const unsigned int c = tile_ch & 0xffffu;
        chr_ok = c < (unsigned int)P.nchr;
        rng = S.chr_rng[chr_ok ? c : 0u];
    }
    unsigned int codes = 0;
    // not unrolled: four copies of the group body are 58 KB of instructions, and without a barrier the warps of an SM
    // spread over all of it (ncu on the unrolled form: 1.9 warps per issue slot waiting for an instruction fetch).
    // Measured and dropped (B200, 300 M contacts, this loop at 2.64 ms): the loop rotated so that the closed forms and
    // stores of group h - 1 run between the gathers of group h and their first use (2.97 ms: the carried state costs
    // registers, 24 B spilled, one more trip); 48 registers and 5 CTAs per SM (3.5 ms before, 4.2 ms
    // after the instruction diet -- the spills weigh more the fewer instructions are left; the switch is gone).
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
        const int l0 = (h * kFrontThreads + tid) * 2;
        const unsigned int pr = pair0 + h * kFrontThreads, i0 = 2u * pr;
        StreamRegs cur;
        if (full) {
            cur = nx;
            if (h < 3 || next_pair0 != 0xffffffffu) stream_issue<PRE>(P, h < 3 ? pr + kFrontThreads : next_pair0, nx);
        }
        int cc[2];
        PvalClass cls[2];
        double prior[2], pv[2], e[2];
        bool use_inter[2];
        if (PRE) {
            unsigned int pc[2];
            double b12[2];
            if (full) {
                cc[0] = cur.a.x; cc[1] = cur.a.y;
                pc[0] = (unsigned int)cur.b.x; pc[1] = (unsigned int)cur.b.y;
                b12[0] = cur.d.x; b12[1] = cur.d.y;
            } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const unsigned int i = i0 + k;
                    const bool ok = i < n32;
                    cc[k] = ok ? reinterpret_cast<const int *>(P.cnt)[i] : 0;
                    pc[k] = ok ? P.pre_code[i] : kPreNotScored;
                    b12[k] = ok ? P.pre_b12[i] : 1.0;
                }
            }
            double gtv[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // the one gather that is left (NaN where the line uses no table value)
                const unsigned int slot = pc[k] & kPreSlotMask;
                const bool want = (pc[k] != kPreNotScored) & !(pc[k] & kPreInter) & (slot < F.D32);  // D32 = 0: no table
                gtv[k] = __ldg(want ? P.lut + slot : reinterpret_cast<const double *>(g_front_consts) + 1);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                FrontFlags f;
                f.scored = pc[k] != kPreNotScored;
                f.intra_path = !(pc[k] & kPreInter);
                f.b_ok = (pc[k] & kPreBok) != 0;
                cls[k] = front_score(P, F, f, cc[k], b12[k], gtv[k], pv[k], e[k], prior[k], use_inter[k]);
            }
        } else {
            int m1[2], m2[2];
            unsigned int ch[2];
            if (full) {
                m1[0] = cur.a.x; m1[1] = cur.a.y;
                m2[0] = cur.b.x; m2[1] = cur.b.y;
                cc[0] = cur.c.x; cc[1] = cur.c.y;
                if (!INTRA && P.chrs != nullptr) {
                    const int2 ah = ldg_stream2(reinterpret_cast<const int2 *>(P.chrs) + pr);
                    ch[0] = (unsigned int)ah.x; ch[1] = (unsigned int)ah.y;
                } else {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        ch[k] = (INTRA || tile_one_run) ? tile_ch
                                                        : run_lookup2(P.run_start, P.run_val, P.nruns, P.line_base + (i0 + k));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const unsigned int i = i0 + k;
                    const bool ok = i < n32;
                    m1[k] = ok ? reinterpret_cast<const int *>(P.mid1)[i] : 0;
                    m2[k] = ok ? reinterpret_cast<const int *>(P.mid2)[i] : 0;
                    cc[k] = ok ? reinterpret_cast<const int *>(P.cnt)[i] : 0;
                    ch[k] = 0x00010000u;  // padding: an inter line
                    if (ok)
                        ch[k] = P.chrs != nullptr ? reinterpret_cast<const unsigned int *>(P.chrs)[i]
                                                  : (tile_one_run ? tile_ch : run_lookup2(P.run_start, P.run_val, P.nruns, P.line_base + i));
                }
            }
            double gb1[2], gb2[2], gtv[2];
            unsigned int dd[2];
            const double *a1[2], *a2[2], *at[2];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                front_addr2<HAS_BIAS, REGULAR, INTRA>(P, F, S.chr_rng, rng32, rng, chr_ok, m1[k], m2[k], ch[k], a1[k], a2[k],
                                                      at[k], dd[k]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                gb1[k] = 1.0;
                gb2[k] = 1.0;
                if (HAS_BIAS) {
                    if (rng32) {  // uniform
                        gb1[k] = __ldg(a1[k]);
                        gb2[k] = __ldg(a2[k]);
                    } else {
                        gb1[k] = bias_lookup_general(P, ch[k] & 0xffffu, m1[k]);
                        gb2[k] = bias_lookup_general(P, ch[k] >> 16, m2[k]);
                    }
                }
                gtv[k] = __ldg(at[k]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool in_file = full || i0 + k < n32;
                cls[k] = front_prepare<INTRA>(P, F, dd[k], cc[k], ch[k], in_file, gb1[k], gb2[k], gtv[k], pv[k], e[k], prior[k],
                                              use_inter[k]);
            }
        }
        codes |= front2_finish_pair(P, S, i0, l0, full, cc, cls, prior, use_inter, pv, e, flagged) << (4 * h);
    }
    return codes;
}

// list positions for the items of one warp and their stores (see the note above)
__device__ __forceinline__ void front2_append(const ListsWs &W, const Front2Smem &S, unsigned int codes, unsigned int base,
                                              int tid, int lane) {
    unsigned int mine = 0;  // continued fractions | tail sums << 16 of this thread (<= 8 each)
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) {
        const unsigned int c = (codes >> (2 * s8)) & 3u;
        mine += (c == kClsCf ? 1u : 0u) + (c == kClsTail ? 0x10000u : 0u);
    }
    unsigned int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const unsigned int tot = __shfl_sync(0xffffffffu, inc, 31);  // <= 256 per half
    if (tot == 0) return;
    unsigned long long b = 0;
    if (lane == 31) b = atomicAdd(W.ctr, (unsigned long long)(tot & 0xffffu) | ((unsigned long long)(tot >> 16) << 32));
    b = __shfl_sync(0xffffffffu, b, 31);
    if (mine == 0) return;
    const unsigned int ex = inc - mine;
    long long oCf = (long long)(b & 0xffffffffull) + (ex & 0xffffu);
    long long oTail = W.cap - 1 - ((long long)(b >> 32) + (ex >> 16));
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) {
        const unsigned int c = (codes >> (2 * s8)) & 3u;
        if (c == kClsCf || c == kClsTail) {
            const int li = ((s8 >> 1) * kFrontThreads + tid) * 2 + (s8 & 1);
            WorkItem it;
            it.x = S.x[li];
            it.idx = base + (unsigned int)li;
            it.cnt = S.cnt[li];
            if (c == kClsCf)
                W.items[oCf++] = it;
            else
                W.items[oTail--] = it;
        }
    }
}

template <bool HAS_BIAS, bool REGULAR, bool PRE>
__global__ void __launch_bounds__(kFrontThreads, 4) pval_front2_kernel(const PvalParams P, const FrontConst F, const ListsWs W) {
    extern __shared__ __align__(16) unsigned char front_smem[];
    Front2Smem &S = *reinterpret_cast<Front2Smem *>(front_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    bool rng32 = false;
    if (HAS_BIAS && !PRE) {
        rng32 = !P.bias_sparse && P.nchr <= kChrSmem && P.chr_off[P.nchr] < 0x7fffffffll;
        if (rng32)
            for (int c = tid; c < P.nchr; c += kFrontThreads) S.chr_rng[c] = make_int2((int)P.chr_off[c], (int)P.chr_off[c + 1]);
    }
    const bool runs = !PRE && P.chrs == nullptr;
    __syncthreads();  // the only one: nothing in the tile loop is exchanged between warps
    const unsigned int ntiles = (unsigned int)((P.n + kFrontTile - 1) / kFrontTile);  // n < 2^32
    const unsigned int nfull = (unsigned int)(P.n / kFrontTile);                        // tiles [0, nfull) are full
    unsigned int flagged = 0;
    int run = 0;  // the tiles of this CTA only move forward, so does its position in the run table
    StreamRegs nx;
    nx.a = nx.b = nx.c = make_int2(0, 0);
    nx.d = make_double2(0.0, 0.0);
    if (blockIdx.x < nfull) stream_issue<PRE>(P, blockIdx.x * (kFrontTile / 2) + tid, nx);

    for (unsigned int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned int base = tile * kFrontTile;  // < 2^32
        const bool full = tile < nfull;
        const unsigned int tnext = tile + gridDim.x;
        const unsigned int pair0 = tile * (kFrontTile / 2) + tid;
        const unsigned int next_pair0 = tnext < nfull ? tnext * (kFrontTile / 2) + tid : 0xffffffffu;
        bool one_run = false;
        unsigned int tile_ch = 0;
        if (runs) {
            const long long g0 = P.line_base + base;
            long long next_start = __ldg(P.run_start + run + 1);
            while (next_start <= g0) next_start = __ldg(P.run_start + (++run) + 1);
            const long long g1 = P.line_base + (full ? (long long)base + kFrontTile : P.n);
            one_run = next_start >= g1;
            tile_ch = __ldg(P.run_val + run);
        }
        unsigned int codes;
        if (PRE)
            codes = front2_tile<false, true, false, true>(P, F, S, rng32, pair0, full, one_run, tile_ch, nx, next_pair0, flagged);
        else if (one_run && full && (tile_ch & 0xffffu) == (tile_ch >> 16) && P.mode != FHC_MODE_INTER_ONLY)
            codes = front2_tile<HAS_BIAS, REGULAR, true, false>(P, F, S, rng32, pair0, full, one_run, tile_ch, nx, next_pair0,
                                                                flagged);
        else
            codes = front2_tile<HAS_BIAS, REGULAR, false, false>(P, F, S, rng32, pair0, full, one_run, tile_ch, nx, next_pair0,
                                                                 flagged);
        front2_append(W, S, codes, base, tid, lane);
    }
    if (P.outl != nullptr) {
        const unsigned long long f = warp_sum((unsigned long long)flagged);
        if (lane == 0 && f) atomicAdd(P.outl_stats, f);
    }
}

// ---- pre-pass ---------------------------------------------------------------------------------------------------------
// Everything of the front kernel that does not depend on the spline table -- the two bias gathers, their product and
// window test, the class of the line (fithic/fithic.py:1057-1115) and its distance slot -- for every contact, 12 bytes out.
// It is launched right after K1 and runs while the host bins and fits (the GPU idles there otherwise); later spline passes
// of the same run reuse its output.
struct PreSmem {
    long long run_start[FHC_MAX_CHR_RUNS + 1];
    unsigned int run_val[FHC_MAX_CHR_RUNS];
    int2 chr_rng[kChrSmem];
};

template <bool HAS_BIAS, bool REGULAR, bool INTRA>
__device__ __forceinline__ void prepass_pair(const PvalParams &P, const FrontConst &F, const PreSmem &S, bool rng32, int2 rng,
                                             bool chr_ok, const int *m1, const int *m2, const unsigned int *ch, bool in_file0,
                                             bool in_file1, unsigned int *code, double *b12) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const unsigned int c1 = ch[k] & 0xffffu, c2 = ch[k] >> 16;
        const unsigned int d = m1[k] > m2[k] ? (unsigned int)m1[k] - (unsigned int)m2[k] : (unsigned int)m2[k] - (unsigned int)m1[k];
        double b1 = 1.0, b2 = 1.0;
        if (HAS_BIAS) {
            if (rng32) {
                if (INTRA) {
                    b1 = bias_lookup_rng<REGULAR>(P, F, rng, chr_ok, m1[k]);
                    b2 = bias_lookup_rng<REGULAR>(P, F, rng, chr_ok, m2[k]);
                } else {
                    b1 = bias_lookup_sel<REGULAR>(P, F, S.chr_rng, c1, m1[k]);
                    b2 = bias_lookup_sel<REGULAR>(P, F, S.chr_rng, c2, m2[k]);
                }
            } else {
                b1 = bias_lookup(P, c1, m1[k]);
                b2 = bias_lookup(P, c2, m2[k]);
            }
        }
        const FrontFlags f = front_flags<INTRA>(P, F, d, ch[k], k == 0 ? in_file0 : in_file1, b1, b2);
        unsigned int slot = d < 0x80000000u ? fastdiv31(d, P, F) : fastdiv(d, P.res);
        slot = slot < kPreSlotMask ? slot : kPreSlotMask;
        code[k] = !f.scored ? kPreNotScored : ((f.b_ok ? kPreBok : 0u) | (f.intra_path ? slot : kPreInter));
        b12[k] = __dmul_rn(b1, b2);
    }
}

template <bool HAS_BIAS, bool REGULAR>
__global__ void __launch_bounds__(256, 4) pval_prepass_kernel(const PvalParams P, const FrontConst F,
                                                              unsigned int *__restrict__ code_out, double *__restrict__ b12_out) {
    __shared__ PreSmem S;
    const int tid = threadIdx.x;
    bool rng32 = false;
    if (HAS_BIAS) {
        rng32 = !P.bias_sparse && P.nchr <= kChrSmem && P.chr_off[P.nchr] < 0x7fffffffll;
        if (rng32)
            for (int c = tid; c < P.nchr; c += 256) S.chr_rng[c] = make_int2((int)P.chr_off[c], (int)P.chr_off[c + 1]);
    }
    const bool runs = P.chrs == nullptr;
    if (runs) {
        for (int r = tid; r <= P.nruns; r += 256) S.run_start[r] = P.run_start[r];
        for (int r = tid; r < P.nruns; r += 256) S.run_val[r] = P.run_val[r];
    }
    __syncthreads();
    const int *m1s = reinterpret_cast<const int *>(P.mid1), *m2s = reinterpret_cast<const int *>(P.mid2);
    const unsigned int *hs = reinterpret_cast<const unsigned int *>(P.chrs);
    const long long npairs = (P.n + 1) >> 1;
    int run = 0;
    for (long long g = (long long)blockIdx.x * 256 + tid; g < npairs; g += (long long)gridDim.x * 256) {
        const long long i0 = g << 1;
        const bool two = i0 + 1 < P.n;
        int m1[2], m2[2];
        unsigned int ch[2];
        bool one_run = false;
        if (two) {
            const int2 a1 = ldg_stream2(reinterpret_cast<const int2 *>(P.mid1) + g);
            const int2 a2 = ldg_stream2(reinterpret_cast<const int2 *>(P.mid2) + g);
            m1[0] = a1.x; m1[1] = a1.y;
            m2[0] = a2.x; m2[1] = a2.y;
        } else {
            m1[0] = m1s[i0]; m2[0] = m2s[i0];
            m1[1] = 0; m2[1] = 0;
        }
        if (!runs) {
            ch[0] = hs[i0];
            ch[1] = two ? hs[i0 + 1] : 0x00010000u;
        } else {
            const long long gl = P.line_base + i0;
            while (S.run_start[run + 1] <= gl) ++run;
            ch[0] = S.run_val[run];
            one_run = S.run_start[run + 1] > gl + 1;
            ch[1] = one_run ? ch[0] : (two ? S.run_val[run + 1 < P.nruns ? run + 1 : run] : 0x00010000u);
            if (!one_run && two) {  // (runs may be empty in theory: find the run of the second line properly)
                int r2 = run;
                while (S.run_start[r2 + 1] <= gl + 1) ++r2;
                ch[1] = S.run_val[r2];
            }
        }
        unsigned int code[2];
        double b12[2];
        const bool intra = runs && one_run && (ch[0] & 0xffffu) == (ch[0] >> 16) && P.mode != FHC_MODE_INTER_ONLY;
        if (intra) {
            int2 rng = make_int2(0, 0);
            bool chr_ok = false;
            if (HAS_BIAS && rng32) {
                const unsigned int c = ch[0] & 0xffffu;
                chr_ok = c < (unsigned int)P.nchr;
                rng = S.chr_rng[chr_ok ? c : 0u];
            }
            prepass_pair<HAS_BIAS, REGULAR, true>(P, F, S, rng32, rng, chr_ok, m1, m2, ch, true, two, code, b12);
        } else {
            prepass_pair<HAS_BIAS, REGULAR, false>(P, F, S, rng32, make_int2(0, 0), false, m1, m2, ch, true, two, code, b12);
        }
        if (two) {
            reinterpret_cast<uint2 *>(code_out)[g] = make_uint2(code[0], code[1]);
            reinterpret_cast<double2 *>(b12_out)[g] = make_double2(b12[0], b12[1]);
        } else {
            code_out[i0] = code[0];
            b12_out[i0] = b12[0];
        }
    }
}

// ---- iterate ----------------------------------------------------------------------------------------------------------
// One list, one kind of recurrence.  A warp claims a chunk of kIterChunk list positions with one global atomic, loads the
// chunk coalesced and prepares every item's set-up (the reciprocal behind z or cN) with all 32 lanes into its own slice of
// shared memory; a lane whose item has converged stores numerator/denominator and starts the next prepared item of the
// chunk at once, so the lanes stay busy until the list ends.  (First version: set-up inside the refill and the item read
// from global memory there -- 15 of 32 lanes active, 37 % of the stall samples on that load.)
struct __align__(16) Staged {
    double zc;  // continued fraction: z = x or x / (1 - x); tail sum: cN = (1 - x) / (x N)
    int cnt;    // count | inter << 31
    int aux;    // bit 0: continued fraction by cephes incbd instead of incbcf; bits 1...: position of the item in its chunk
};

// Tail sums of a few terms stay out of the queue (tail_is_short, cephes_dev.cuh): while a chunk is staged -- all lanes
// busy, ~10 instructions per item -- the items that remain are compacted to the front of the warp's slice.  (ncu before
// that, bench input with counts around their expectation: 12.4 M trips through the tail loop against 1.6 M through the
// continued fractions, 7 lanes refilled per trip, almost every item finished after one or two terms -- 80 % of the
// kernel was spent handing out items.)
template <bool TAIL>
__device__ __forceinline__ void iterate_list(const PvalParams &P, const ListsWs &W, unsigned long long total,
                                             unsigned long long *cursor, int lane, Staged *stage) {
    if (total == 0) return;
    CfState st;
    long long pos = -1;
    unsigned long long base = 0;
    unsigned int lo = 0, hi = 0;  // staged items of the chunk not handed out yet: stage[lo, hi)
    bool drained = false;
    const unsigned int lt = (1u << lane) - 1u;
    // the next chunk is claimed one chunk ahead (the round trip of the atomic hides behind the staging of the current one;
    // claims beyond the end of the list only move the cursor)
    unsigned long long claimed = 0;
    if (lane == 0) claimed = atomicAdd(cursor, (unsigned long long)kIterChunk);
    while (true) {
        const unsigned int m = __ballot_sync(0xffffffffu, pos < 0);
        if (m != 0 && !drained) {
            while (lo >= hi && !drained) {  // take chunks until one holds work or the list ends
                const unsigned long long b = __shfl_sync(0xffffffffu, claimed, 0);
                if (lane == 0 && b < total) claimed = atomicAdd(cursor, (unsigned long long)kIterChunk);
                base = b;
                lo = 0;
                hi = 0;
                const unsigned int nchunk =
                    b < total ? (unsigned int)(total - b < (unsigned long long)kIterChunk ? total - b : kIterChunk) : 0u;
                if (nchunk == 0) {
                    drained = true;
                    break;
                }
                __syncwarp();  // every lane has read what it needed from the previous chunk
                WorkItem its[kIterChunk / 32];  // all loads of the chunk in flight before the first one is looked at
#pragma unroll
                for (int r = 0; r < kIterChunk / 32; ++r) {
                    const unsigned int i = r * 32 + lane;
                    const long long ip = TAIL ? (W.cap - 1 - (long long)(base + i)) : (long long)(base + i);
                    its[r].x = 0.0;
                    its[r].idx = 0u;
                    its[r].cnt = 0;
                    if (i < nchunk) its[r] = W.items[ip];
                }
#pragma unroll
                for (int r = 0; r < kIterChunk / 32; ++r) {
                    const unsigned int i = r * 32 + lane;
                    bool keep = false;
                    Staged sg;
                    sg.zc = 0.0;
                    sg.cnt = 0;
                    sg.aux = 0;
                    if (i < nchunk) {
                        const WorkItem it = its[r];
                        keep = !TAIL || !tail_is_short(it.cnt & 0x7fffffff);
                        if (keep) {
                            const bool ui = it.cnt < 0;
                            const double dN = (double)(ui ? P.N_inter : P.N_intra);
                            sg.cnt = it.cnt;
                            if (TAIL) {
                                sg.zc = tail_cn(dN, it.x, __dsub_rn(1.0, it.x));
                                sg.aux = (int)(i << 1);
                            } else {
                                const double aa = (double)(it.cnt & 0x7fffffff);
                                const bool use_d = cf_uses_d(aa, dN - aa + 1.0, it.x);
                                sg.zc = cf_z(it.x, use_d);
                                sg.aux = (int)(i << 1) | (use_d ? 1 : 0);
                            }
                        }
                    }
                    const unsigned int bm = __ballot_sync(0xffffffffu, keep);
                    if (keep) stage[hi + (unsigned int)__popc(bm & lt)] = sg;
                    hi += (unsigned int)__popc(bm);
                }
                __syncwarp();
            }
            if (pos < 0 && !drained) {
                const unsigned int k = lo + (unsigned int)__popc(m & lt);
                if (k < hi) {
                    const Staged sg = stage[k];
                    const unsigned int off = (unsigned int)sg.aux >> 1;
                    pos = TAIL ? (W.cap - 1 - (long long)(base + off)) : (long long)(base + off);
                    const bool ui = sg.cnt < 0;
                    const double dN = (double)(ui ? P.N_inter : P.N_intra);
                    const double aa = (double)(sg.cnt & 0x7fffffff);
                    if (TAIL)
                        tail_fwd_load(st, aa, dN, ui ? P.invN_inter : P.invN_intra, sg.zc);
                    else
                        cf_load(st, aa, dN - aa + 1.0, sg.zc, (sg.aux & 1) != 0);
                }
            }
            const unsigned int adv = lo + (unsigned int)__popc(m);
            lo = adv < hi ? adv : hi;
        }
        if (__ballot_sync(0xffffffffu, pos >= 0) == 0) {
            if (drained) break;
            continue;
        }
        if (pos >= 0) {
            const bool done = TAIL ? tail_fwd_step(st) : cf_step(st);
            if (done) {
                W.pq[pos] = make_double2(st.pkm1, st.qkm1);
                pos = -1;
            }
        }
    }
}

__global__ void __launch_bounds__(kIterThreads, 4) pval_iterate_kernel(const PvalParams P, const ListsWs W) {
    __shared__ Staged stage_all[kIterThreads / 32][kIterChunk];
    const int lane = threadIdx.x & 31;
    Staged *stage = stage_all[threadIdx.x >> 5];
    const unsigned long long nCf = W.ctr[0] & 0xffffffffull, nTail = W.ctr[0] >> 32;
    iterate_list<false>(P, W, nCf, W.ctr + 2, lane, stage);
    iterate_list<true>(P, W, nTail, W.ctr + 3, lane, stage);
}

// ---- finish -----------------------------------------------------------------------------------------------------------
// aux[c] = {lbeta(c, N-c+1) + log(c), lbeta(c, N-c+1) + log(N-c+1)} from the lbeta table of the run
__global__ void lbeta_aux_kernel(const double *__restrict__ tab, long long ntab, int N, double2 *__restrict__ aux) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ntab) return;
    double2 v = make_double2(NAN, NAN);
    if (c >= 1 && c <= (long long)N) {
        const double lb = tab[c];
        v = make_double2(lb + log((double)c), lb + log((double)((long long)N - c + 1)));
    }
    aux[c] = v;
}

// Two items per thread and iteration: their four 16-byte loads are issued before the first logarithm starts (the kernel
// waits on memory, not on arithmetic: two thirds of its stall samples sat on these loads with one item per thread).
// Measured and dropped: the four loads of the NEXT iteration issued before the current two items are worked on (1.40 ms
// against 1.25 ms: 16 more live registers at the 80 the kernel has, 200 B spilled).
template <int kMinCtas>
__global__ void __launch_bounds__(kFinishThreads, kMinCtas) pval_finish_kernel(const PvalParams P, const ListsWs W) {
    const unsigned long long nCf = W.ctr[0] & 0xffffffffull, nTail = W.ctr[0] >> 32;
    const unsigned long long total = nCf + nTail;
    const unsigned long long stride = (unsigned long long)gridDim.x * kFinishThreads;
    unsigned int flagged = 0;
    for (unsigned long long k0 = (unsigned long long)blockIdx.x * kFinishThreads + threadIdx.x; k0 < total; k0 += 2 * stride) {
        WorkItem it[2];
        double2 pq[2];
        bool tail[2], live[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned long long k = k0 + j * stride;
            live[j] = k < total;
            tail[j] = k >= nCf;
            const long long pos = live[j] ? (tail[j] ? (W.cap - 1 - (long long)(k - nCf)) : (long long)k) : 0;
            it[j] = W.items[pos];
            pq[j] = W.pq[pos];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!live[j]) continue;
            const bool ui = it[j].cnt < 0;
            const int c = it[j].cnt & 0x7fffffff;
            const int N = ui ? P.N_inter : P.N_intra;
            const double aa = (double)c, bb = (double)((long long)N - c + 1);
            const double2 *aux = ui ? W.aux_inter : W.aux_intra;
            const long long ntab = ui ? P.ntab_inter : P.ntab_intra;
            double2 a;
            if (aux != nullptr && c < ntab) {
                a = __ldg(aux + c);
            } else {
                const double lb = lbeta_cephes(aa, bb);
                a = make_double2(lb + log(aa), lb + log(bb));
            }
            double w;
            if (tail[j] && tail_is_short(c)) {  // the few terms of a short tail sum, as pval_iterate_kernel would run them
                const double dN = (double)N;
                double q;
                const double pn = tail_short_sum(aa, dN, ui ? P.invN_inter : P.invN_intra,
                                                 tail_cn(dN, it[j].x, __dsub_rn(1.0, it[j].x)), &q);
                w = pn / q;
            } else {
                w = pq[j].x / pq[j].y;
            }
            const double p = incbet_finish_folded(tail[j], aa, bb, it[j].x, a.x, a.y, w);
            P.p[it[j].idx] = p;
            if (P.outl != nullptr) outlier_mark(P, (long long)it[j].idx, p, flagged);
        }
    }
    if (P.outl != nullptr) {
        const unsigned long long f = warp_sum((unsigned long long)flagged);
        if ((threadIdx.x & 31) == 0 && f) atomicAdd(P.outl_stats, f);
    }
}

// ---- host -------------------------------------------------------------------------------------------------------------
static FrontConst make_front_const(const PvalParams &P) {
    FrontConst F;
    F.nothing_in_range = P.Llo > 0xffffffffll;
    F.Llo = (unsigned int)(P.Llo > 0xffffffffll ? 0xffffffffll : P.Llo);
    F.Uhi = (unsigned int)(P.Uhi > 0xffffffffll ? 0xffffffffll : P.Uhi);
    F.dN_intra = (double)P.N_intra;
    F.dN_inter = (double)P.N_inter;
    F.dNp1_intra = (double)P.N_intra + 1.0;
    F.dNp1_inter = (double)P.N_inter + 1.0;
    const unsigned int d = P.res.d;
    unsigned int l = 0;
    while ((1ull << l) < d) ++l;  // ceil(log2 d)
    F.div_sh = l ? l - 1 : 0;
    F.div_m = d > 1 ? (unsigned int)(((1ull << (31 + l)) + d - 1) / d) : 0u;
    F.D32 = P.lut == nullptr ? 0u : (unsigned int)(P.D > 0xffffffffll ? 0xffffffffll : P.D);
    F.half = d >> 1;
    return F;
}

int pvalues_prepass_launch(const PvalParams &P, unsigned int *code, double *b12, cudaStream_t st) {
    const FrontConst F = make_front_const(P);
    long long blocks = (((P.n + 1) >> 1) + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    if (!P.bias)
        pval_prepass_kernel<false, true><<<(unsigned int)blocks, 256, 0, st>>>(P, F, code, b12);
    else if (P.bias_mid == nullptr && !P.bias_sparse)
        pval_prepass_kernel<true, true><<<(unsigned int)blocks, 256, 0, st>>>(P, F, code, b12);
    else
        pval_prepass_kernel<true, false><<<(unsigned int)blocks, 256, 0, st>>>(P, F, code, b12);
    FHC_LAUNCH_CHECK("pval_prepass_kernel");
    return FHC_OK;
}

static size_t lists_align(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t lists_layout(long long n, long long ntab_intra, long long ntab_inter, char *base, ListsWs *ws) {
    const size_t nn = (size_t)(n > 0 ? n : 1);
    size_t off = 0;
    if (ws) ws->ctr = reinterpret_cast<unsigned long long *>(base + off);
    off += 256;
    if (ws) ws->items = reinterpret_cast<WorkItem *>(base + off);
    off += lists_align(nn * sizeof(WorkItem));
    if (ws) ws->pq = reinterpret_cast<double2 *>(base + off);
    off += lists_align(nn * sizeof(double2));
    if (ws) ws->aux_intra = ntab_intra > 0 ? reinterpret_cast<double2 *>(base + off) : nullptr;
    off += lists_align((size_t)(ntab_intra > 0 ? ntab_intra : 0) * sizeof(double2));
    if (ws) ws->aux_inter = ntab_inter > 0 ? reinterpret_cast<double2 *>(base + off) : nullptr;
    off += lists_align((size_t)(ntab_inter > 0 ? ntab_inter : 0) * sizeof(double2));
    if (ws) ws->cap = (long long)nn;
    return off;
}

size_t pvalues_lists_workspace_bytes(long long n, long long ntab) {
    // the split of ntab between the two tables is not known here: reserve it for both
    return lists_layout(n, ntab, ntab, nullptr, nullptr);
}

int pvalues_lists_launch(const PvalParams &P, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const long long n = P.n;
    FHC_REQUIRE(n < (1ll << 32), FHC_E_INVALID, "fhc_pvalues: the work-list path takes at most 2^32 - 1 contacts per call");
    const size_t need = lists_layout(n, P.ntab_intra, P.ntab_inter, nullptr, nullptr);
    FHC_REQUIRE(workspace_bytes >= need, FHC_E_WORKSPACE, "fhc_pvalues: workspace of %zu bytes, need %zu", workspace_bytes,
                need);
    FHC_REQUIRE(aligned16(workspace), FHC_E_INVALID, "fhc_pvalues: workspace must be 16-byte aligned");
    ListsWs W;
    lists_layout(n, P.ntab_intra, P.ntab_inter, reinterpret_cast<char *>(workspace), &W);
    FHC_CUDA(cudaMemsetAsync(W.ctr, 0, 256, st));
    if (W.aux_intra != nullptr) {
        lbeta_aux_kernel<<<(unsigned int)((P.ntab_intra + 127) / 128), 128, 0, st>>>(P.lbeta_intra, P.ntab_intra, P.N_intra,
                                                                                  W.aux_intra);
        FHC_LAUNCH_CHECK("lbeta_aux_kernel");
    }
    if (W.aux_inter != nullptr) {
        lbeta_aux_kernel<<<(unsigned int)((P.ntab_inter + 127) / 128), 128, 0, st>>>(P.lbeta_inter, P.ntab_inter, P.N_inter,
                                                                                  W.aux_inter);
        FHC_LAUNCH_CHECK("lbeta_aux_kernel");
    }
    const FrontConst F = make_front_const(P);
    long long tiles = (n + kFrontTile - 1) / kFrontTile;
    long long blocks = tiles;
    // Default: the second version of the front kernel (streamed words loaded one group ahead, list positions per warp,
    // no barrier in the tile loop).  FHC_PVAL_FRONT=v1 selects the first version (two contacts per load, CTA-wide scan and
    // two barriers per tile) for comparison.  Measured on the first version and dropped (B200, 300 M contacts in random
    // order): four contacts per load in 80 registers and 3 CTAs per SM (5.25 ms against 4.63 ms), the same squeezed into
    // 64 registers with spills (5.40 ms); the contact arrays of the next tile staged in shared memory by cp.async with the
    // gathers issued one pair ahead (6.65 ms: the extra barriers cost more than the staging saves); an L2 prefetch of the
    // CTA's next tile (4.02 ms with, 4.04 ms without).
    const char *fv = getenv("FHC_PVAL_FRONT");
    // (the second version drops the 64-bit division for distances of 2^31 and more -- one mid point negative --, which is
    // only right while such a distance lies beyond the table: a table that reaches further goes to the first version)
    const bool wide_table = (unsigned long long)F.D32 * P.res.d > 0x80000000ull;
    const bool v1 = (fv && fv[0] == 'v' && fv[1] == '1') || wide_table;
    if (blocks > (long long)kNumSMs * 4) blocks = (long long)kNumSMs * 4;
#define FHC_FRONT_V(B, R, PRE)                                                                                         \
    do {                                                                                                               \
        if (v1)                                                                                                        \
            pval_front_kernel<B, R, 4, 2, PRE><<<(unsigned int)blocks, kFrontThreads, sizeof(FrontSmem), st>>>(P, F, W);  \
        else                                                                                                           \
            pval_front2_kernel<B, R, PRE><<<(unsigned int)blocks, kFrontThreads, sizeof(Front2Smem), st>>>(P, F, W); \
    } while (0)
    if (P.pre_code != nullptr)
        FHC_FRONT_V(false, true, true);
    else if (!P.bias)
        FHC_FRONT_V(false, true, false);
    else if (P.bias_mid == nullptr && !P.bias_sparse)
        FHC_FRONT_V(true, true, false);
    else
        FHC_FRONT_V(true, false, false);
#undef FHC_FRONT_V
    FHC_LAUNCH_CHECK(v1 ? "pval_front_kernel" : "pval_front2_kernel");
    // the list lengths are only known on the device: both follow-up kernels are persistent and read them there
    long long iblocks = (n + kIterThreads * 4 - 1) / (kIterThreads * 4);
    if (iblocks > (long long)kNumSMs * 4) iblocks = (long long)kNumSMs * 4;
    pval_iterate_kernel<<<(unsigned int)iblocks, kIterThreads, 0, st>>>(P, W);
    FHC_LAUNCH_CHECK("pval_iterate_kernel");
    long long fblocks = (n + kFinishThreads * 4 - 1) / (kFinishThreads * 4);
    if (fblocks > (long long)kNumSMs * 8) fblocks = (long long)kNumSMs * 8;
    const char *nocc = getenv("FHC_PVAL_FINISH_OCC");  // 3 (default): 80 registers; 4: 64 registers with ~150 B spilled
    if (nocc && nocc[0] == '4')                        // (B200, 300 M contacts: 1.61 ms against 1.83 ms)
        pval_finish_kernel<4><<<(unsigned int)fblocks, kFinishThreads, 0, st>>>(P, W);
    else
        pval_finish_kernel<3><<<(unsigned int)fblocks, kFinishThreads, 0, st>>>(P, W);
    FHC_LAUNCH_CHECK("pval_finish_kernel");
    return FHC_OK;
}

}  // namespace fhc
