// Parameters and the per-line branch order of fit_Spline (reference fithic/fithic.py:1017-1123), shared by the two
// implementations of K3: the tile-phased single kernel (pvalue.cu) and the work-list pipeline (pvalue_lists.cu).
#pragma once
#include "cephes_dev.cuh"

namespace fhc {

// n / d for 32-bit n by one 64-bit multiply-high: M = ceil(2^64 / d) is exact for every n < 2^32 (d == 1 is flagged)
struct FastDiv {
    unsigned long long M;
    unsigned int d;
};
__device__ __forceinline__ unsigned int fastdiv(unsigned int n, const FastDiv &f) {
    return f.d == 1 ? n : (unsigned int)__umul64hi((unsigned long long)n, f.M);
}

struct PvalParams {
    int mode;  // FHC_MODE_*
    const int4 *mid1, *mid2, *cnt, *chrs;  // chrs == nullptr: the chromosome ids come as runs (work-list pipeline only)
    const long long *run_start;            // run r covers lines [run_start[r], run_start[r + 1]) of the caller's arrays,
    const unsigned int *run_val;           // counted like line_base (this call starts at line_base); all have chrs = run_val[r]
    int nruns;
    long long n;
    const double *bias;
    const int *bias_mid;  // nullptr: every slot holds the locus at mid = slot * res + res / 2 (regular grid)
    int bias_sparse;      // 1: no grid (restriction fragments): bias_mid holds each chromosome's mid points in ascending
                          //    order, a locus is found by binary search
    const long long *chr_off;
    int nchr;
    FastDiv res;
    long long Llo, Uhi;  // effective in-range window (L == -1 -> 0, U == -1 -> max)
    const double *lut;
    long long D;
    int N_intra, N_inter;
    double invN_intra, invN_inter;
    double interChrProb, tL, tU;
    const double *lbeta_intra, *lbeta_inter;
    long long ntab_intra, ntab_inter;
    unsigned char *outl;
    long long line_base;  // index of the first contact of this call in the whole file (for the outlier statistics)
    double outl_thres;
    unsigned long long *outl_stats;
    double *p, *expcc;
    // output of fhc_pvalues_prepass for the same contacts (both or neither; work-list pipeline only): with them the front
    // kernel touches neither mid points, chromosome ids nor the bias table
    const unsigned int *pre_code;
    const double *pre_b12;
};

// bias dictionary lookup of fithic/fithic.py:1026-1054: missing chromosome or mid point -> -1
__device__ __forceinline__ double bias_lookup(const PvalParams &P, unsigned int chr, int mid) {
    if ((int)chr >= P.nchr || mid < 0) return -1.0;
    const long long lo = __ldg(P.chr_off + chr), hi = __ldg(P.chr_off + chr + 1);
    if (P.bias_sparse) {
        long long a = lo, b = hi;  // lower bound of mid in bias_mid[lo, hi)
        while (a < b) {
            const long long m = (a + b) >> 1;
            if (__ldg(P.bias_mid + m) < mid)
                a = m + 1;
            else
                b = m;
        }
        return (a < hi && __ldg(P.bias_mid + a) == mid) ? __ldg(P.bias + a) : -1.0;
    }
    const unsigned int k = fastdiv((unsigned int)mid, P.res);
    const long long s = lo + (long long)k;
    if (s >= hi) return -1.0;
    if (P.bias_mid != nullptr) {
        if (__ldg(P.bias_mid + s) != mid) return -1.0;
    } else if ((unsigned int)mid - k * P.res.d != (P.res.d >> 1)) {
        return -1.0;
    }
    return __ldg(P.bias + s);
}

// The branch order of the reference's per-line loop (fithic/fithic.py:1057-1115) up to the bdtrc call.
// Returns the evaluation class; `p` holds the result when the class is kClsDone.
template <bool HAS_BIAS>
__device__ __forceinline__ PvalClass pval_prepare(const PvalParams &P, int m1, int m2, int c, unsigned int ch, double &p,
                                                  double &e, double &prior, bool &use_inter) {
    const unsigned int c1 = ch & 0xffffu, c2 = ch >> 16;
    const bool inter = c1 != c2;
    long long d = (long long)m1 - (long long)m2;
    d = d < 0 ? -d : d;
    double b1 = 1.0, b2 = 1.0;
    if (HAS_BIAS) {
        b1 = bias_lookup(P, c1, m1);
        b2 = bias_lookup(P, c2, m2);
    }
    const bool interOnly = P.mode == FHC_MODE_INTER_ONLY;
    p = 1.0;
    e = 0.0;
    prior = 0.0;
    use_inter = false;
    if ((b1 < 0.0 || b2 < 0.0) && !inter) return kClsDone;  // discarded locus (:1057-1063)
    int N;
    if (!inter && !interOnly) {
        if (!(d >= P.Llo && d <= P.Uhi)) return kClsDone;  // intraShort / intraLong: p = 1, ExpCC = 0 (:1081-1096)
        const unsigned int slot = fastdiv((unsigned int)d, P.res);  // intraInRange (:1065-1079)
        const double prior0 = ((long long)slot < P.D) ? __ldg(P.lut + slot) : NAN;
        prior = __dmul_rn(prior0, __dmul_rn(b1, b2));
        N = P.N_intra;
    } else if (P.mode != FHC_MODE_INTRA_ONLY) {
        // inter lines, and under interOnly every line that was not discarded (:1098-1108)
        prior = __dmul_rn(P.interChrProb, __dmul_rn(b1, b2));
        N = P.N_inter;
        use_inter = true;
    } else {
        return kClsDone;  // inter line in an intraOnly run (:1110-1115)
    }
    if (b1 >= P.tL && b1 <= P.tU && b2 >= P.tL && b2 <= P.tU) e = __dmul_rn((double)N, prior);
    return bdtrc_classify(c, N, prior, p);
}

__device__ __forceinline__ void outlier_mark(const PvalParams &P, long long i, double p, unsigned int &flagged) {
    if (p < P.outl_thres) {  // NaN compares false (:1215)
        const unsigned char m = P.outl[i];
        const unsigned char m2 = m == 255 ? 255 : m + 1;
        P.outl[i] = m2;
        flagged += 1;
        if (m2 >= 2) atomicMin(P.outl_stats + 1, (unsigned long long)(P.line_base + i));
    }
}

// implemented in pvalue.cu (tile-phased kernel) and pvalue_lists.cu (work-list pipeline)
int pvalues_tile_launch(const PvalParams &P, cudaStream_t st);
size_t pvalues_lists_workspace_bytes(long long n, long long ntab);
int pvalues_lists_launch(const PvalParams &P, void *workspace, size_t workspace_bytes, cudaStream_t st);
int pvalues_prepass_launch(const PvalParams &P, unsigned int *code, double *b12, cudaStream_t st);

}  // namespace fhc
