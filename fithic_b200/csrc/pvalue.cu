// K3 -- fused per-contact significance: classify, bias gather, spline-table lookup, binomial survival, ExpCC.
//
// Replaces the per-line loop of fit_Spline (reference fithic/fithic.py:1017-1123) and scipy.special.bdtrc
// (:1070, :1101).  One thread handles 4 consecutive contacts: four 128-bit streaming loads in (16 B/contact), two
// 128-bit stores out for p and two for ExpCC (16 B/contact).  The bias vector, the distance table and the lbeta table
// are small (<= 5 MB) and stay in L1/L2.  The kernel is FP64-ALU bound (continued fraction + 4 transcendentals per
// contact), not HBM bound.
#include "cephes_dev.cuh"

namespace fhc {

constexpr int kPvalThreads = 256;

struct PvalParams {
    int mode;  // FHC_MODE_*
    const int4 *mid1, *mid2, *cnt, *chrs;
    long long n;
    const double *bias;
    const int *bias_mid;
    const long long *chr_off;
    int nchr;
    unsigned int res;
    long long Llo, Uhi;  // effective in-range window (L == -1 -> 0, U == -1 -> max)
    const double *lut;
    long long D;
    int N_intra, N_inter;
    double interChrProb, tL, tU;
    const double *lbeta_intra, *lbeta_inter;
    long long ntab_intra, ntab_inter;
    unsigned char *outl;
    double outl_thres;
    unsigned long long *outl_stats;
    double *p, *expcc;
};

// bias dictionary lookup of fithic/fithic.py:1026-1054: missing chromosome or mid point -> -1
__device__ __forceinline__ double bias_lookup(const PvalParams &P, unsigned int chr, int mid) {
    if ((int)chr >= P.nchr || mid < 0) return -1.0;
    const long long lo = __ldg(P.chr_off + chr), hi = __ldg(P.chr_off + chr + 1);
    const long long s = lo + (long long)((unsigned int)mid / P.res);
    if (s >= hi) return -1.0;
    if (__ldg(P.bias_mid + s) != mid) return -1.0;
    return __ldg(P.bias + s);
}

template <bool HAS_BIAS>
__device__ __forceinline__ void pval_one(const PvalParams &P, int m1, int m2, int c, unsigned int ch, double &p_out,
                                         double &e_out) {
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
    double p = 1.0, e = 0.0;
    if ((b1 < 0.0 || b2 < 0.0) && !inter) {
        // discarded locus (:1057-1063)
    } else if (!inter && !interOnly) {
        if (d >= P.Llo && d <= P.Uhi) {  // intraInRange (:1065-1079)
            const unsigned int du = (unsigned int)d;
            const unsigned int slot = du / P.res;
            const double prior0 = ((long long)slot < P.D) ? __ldg(P.lut + slot) : NAN;
            const double prior = __dmul_rn(prior0, __dmul_rn(b1, b2));
            p = bdtrc_dev(c, P.N_intra, prior, P.lbeta_intra, P.ntab_intra);
            if (b1 >= P.tL && b1 <= P.tU && b2 >= P.tL && b2 <= P.tU) e = __dmul_rn((double)P.N_intra, prior);
        }
        // intraShort / intraLong: p = 1, ExpCC = 0 (:1081-1096)
    } else if (P.mode != FHC_MODE_INTRA_ONLY) {
        // inter lines, and under interOnly every line that was not discarded (:1098-1108)
        const double prior = __dmul_rn(P.interChrProb, __dmul_rn(b1, b2));
        p = bdtrc_dev(c, P.N_inter, prior, P.lbeta_inter, P.ntab_inter);
        if (b1 >= P.tL && b1 <= P.tU && b2 >= P.tL && b2 <= P.tU) e = __dmul_rn((double)P.N_inter, prior);
    }
    p_out = p;
    e_out = e;
}

__device__ __forceinline__ void outlier_mark(const PvalParams &P, long long i, double p, unsigned int &flagged) {
    if (p < P.outl_thres) {  // NaN compares false (:1215)
        const unsigned char m = P.outl[i];
        const unsigned char m2 = m == 255 ? 255 : m + 1;
        P.outl[i] = m2;
        flagged += 1;
        if (m2 >= 2) atomicMin(P.outl_stats + 1, (unsigned long long)i);
    }
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(kPvalThreads) pvalues_kernel(const PvalParams P) {
    const long long ngroups = P.n >> 2;
    unsigned int flagged = 0;
    for (long long g = (long long)blockIdx.x * kPvalThreads + threadIdx.x; g < ngroups;
         g += (long long)gridDim.x * kPvalThreads) {
        const int4 a1 = ldg_stream(P.mid1 + g), a2 = ldg_stream(P.mid2 + g), ac = ldg_stream(P.cnt + g);
        const int4 ah = ldg_stream(P.chrs + g);
        double p0, p1, p2, p3, e0, e1, e2, e3;
        pval_one<HAS_BIAS>(P, a1.x, a2.x, ac.x, (unsigned int)ah.x, p0, e0);
        pval_one<HAS_BIAS>(P, a1.y, a2.y, ac.y, (unsigned int)ah.y, p1, e1);
        pval_one<HAS_BIAS>(P, a1.z, a2.z, ac.z, (unsigned int)ah.z, p2, e2);
        pval_one<HAS_BIAS>(P, a1.w, a2.w, ac.w, (unsigned int)ah.w, p3, e3);
        double2 *pp = reinterpret_cast<double2 *>(P.p) + 2 * g;
        double2 *ee = reinterpret_cast<double2 *>(P.expcc) + 2 * g;
        __stcs(pp, make_double2(p0, p1));
        __stcs(pp + 1, make_double2(p2, p3));
        __stcs(ee, make_double2(e0, e1));
        __stcs(ee + 1, make_double2(e2, e3));
        if (P.outl != nullptr) {
            outlier_mark(P, 4 * g + 0, p0, flagged);
            outlier_mark(P, 4 * g + 1, p1, flagged);
            outlier_mark(P, 4 * g + 2, p2, flagged);
            outlier_mark(P, 4 * g + 3, p3, flagged);
        }
    }
    // tail (n % 4 contacts)
    if (blockIdx.x == 0 && threadIdx.x < (P.n & 3)) {
        const long long i = (ngroups << 2) + threadIdx.x;
        double p, e;
        pval_one<HAS_BIAS>(P, reinterpret_cast<const int *>(P.mid1)[i], reinterpret_cast<const int *>(P.mid2)[i],
                           reinterpret_cast<const int *>(P.cnt)[i], reinterpret_cast<const unsigned int *>(P.chrs)[i], p,
                           e);
        P.p[i] = p;
        P.expcc[i] = e;
        if (P.outl != nullptr) outlier_mark(P, i, p, flagged);
    }
    if (P.outl != nullptr) {
        const unsigned long long f = warp_sum((unsigned long long)flagged);
        if ((threadIdx.x & 31) == 0 && f) atomicAdd(P.outl_stats, f);
    }
}

__global__ void lbeta_table_kernel(int N, double *tab, long long ntab) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ntab) return;
    double v = NAN;
    if (c >= 1 && c <= (long long)N) v = lbeta_cephes((double)c, (double)((long long)N - c + 1));
    tab[c] = v;
}

__global__ void bdtrc_kernel(const int *__restrict__ km1, int N, const double *__restrict__ prior, long long n,
                             const double *__restrict__ lbeta, long long ntab, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = bdtrc_dev(km1[i] + 1, N, prior[i], lbeta, ntab);
}

}  // namespace fhc

extern "C" int fhc_lbeta_table(int64_t N, double *tab, int64_t ntab, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(tab != nullptr && ntab > 0, FHC_E_INVALID, "fhc_lbeta_table: null table or ntab <= 0");
    FHC_REQUIRE(N >= 0 && N < (1ll << 31), FHC_E_RANGE,
                "fhc_lbeta_table: N = %lld does not fit the int32 that scipy.special.bdtrc truncates n to", (long long)N);
    const int threads = 128;
    const long long blocks = (ntab + threads - 1) / threads;
    lbeta_table_kernel<<<(unsigned int)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>((int)N, tab, ntab);
    FHC_LAUNCH_CHECK("lbeta_table_kernel");
    return FHC_OK;
}

extern "C" int fhc_bdtrc(const int32_t *cnt_minus_1, int64_t N, const double *prior, int64_t n, const double *lbeta,
                         int64_t ntab, double *out, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0, FHC_E_INVALID, "fhc_bdtrc: n < 0");
    FHC_REQUIRE(N >= 0 && N < (1ll << 31), FHC_E_RANGE, "fhc_bdtrc: N = %lld outside int32", (long long)N);
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(cnt_minus_1 && prior && out, FHC_E_INVALID, "fhc_bdtrc: null pointer");
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    bdtrc_kernel<<<(unsigned int)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        cnt_minus_1, (int)N, prior, n, lbeta, lbeta ? ntab : 0, out);
    FHC_LAUNCH_CHECK("bdtrc_kernel");
    return FHC_OK;
}

extern "C" int fhc_pvalues(int32_t mode, const int32_t *mid1, const int32_t *mid2, const int32_t *cnt,
                           const uint32_t *chrs, int64_t n, const double *bias, const int32_t *bias_mid,
                           const int64_t *chr_off, int32_t nchr, int32_t res, int64_t L, int64_t U, const double *lut,
                           int64_t D, int64_t N_intra, int64_t N_inter, double interChrProb, double tL, double tU,
                           const double *lbeta_intra, int64_t ntab_intra, const double *lbeta_inter, int64_t ntab_inter,
                           uint8_t *outl, double outl_thres, uint64_t *outl_stats, double *p, double *expcc,
                           void *stream) {
    using namespace fhc;
    FHC_REQUIRE(mode == FHC_MODE_INTRA_ONLY || mode == FHC_MODE_INTER_ONLY || mode == FHC_MODE_ALL, FHC_E_INVALID,
                "fhc_pvalues: unknown mode %d", mode);
    FHC_REQUIRE(n >= 0 && res > 0, FHC_E_INVALID, "fhc_pvalues: need n >= 0 and res > 0");
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(mid1 && mid2 && cnt && chrs && p && expcc, FHC_E_INVALID, "fhc_pvalues: null pointer");
    FHC_REQUIRE(aligned16(mid1) && aligned16(mid2) && aligned16(cnt) && aligned16(chrs) && aligned16(p) && aligned16(expcc),
                FHC_E_INVALID, "fhc_pvalues: contact and output arrays must be 16-byte aligned");
    FHC_REQUIRE(mode == FHC_MODE_INTER_ONLY || (lut != nullptr && D > 0), FHC_E_INVALID,
                "fhc_pvalues: the distance table is required unless mode is interOnly");
    FHC_REQUIRE(bias == nullptr || (bias_mid && chr_off && nchr > 0), FHC_E_INVALID,
                "fhc_pvalues: bias needs bias_mid, chr_off and nchr");
    FHC_REQUIRE(N_intra >= 0 && N_intra < (1ll << 31) && N_inter >= 0 && N_inter < (1ll << 31), FHC_E_RANGE,
                "fhc_pvalues: N_intra = %lld / N_inter = %lld do not fit the int32 scipy.special.bdtrc truncates n to "
                "(the reference returns NaN or garbage there, SURVEY F5)",
                (long long)N_intra, (long long)N_inter);
    FHC_REQUIRE(outl == nullptr || outl_stats != nullptr, FHC_E_INVALID, "fhc_pvalues: outl needs outl_stats");
    FHC_REQUIRE(L >= -1 && U >= -1, FHC_E_INVALID, "fhc_pvalues: L and U must be >= -1");
    PvalParams P;
    P.mode = mode;
    P.mid1 = reinterpret_cast<const int4 *>(mid1);
    P.mid2 = reinterpret_cast<const int4 *>(mid2);
    P.cnt = reinterpret_cast<const int4 *>(cnt);
    P.chrs = reinterpret_cast<const int4 *>(chrs);
    P.n = n;
    P.bias = bias;
    P.bias_mid = bias_mid;
    P.chr_off = reinterpret_cast<const long long *>(chr_off);
    P.nchr = nchr;
    P.res = (unsigned int)res;
    P.Llo = L < 0 ? 0 : L;
    P.Uhi = U < 0 ? INT64_MAX : U;
    P.lut = lut;
    P.D = lut ? D : 0;
    P.N_intra = (int)N_intra;
    P.N_inter = (int)N_inter;
    P.interChrProb = interChrProb;
    P.tL = tL;
    P.tU = tU;
    P.lbeta_intra = lbeta_intra;
    P.ntab_intra = lbeta_intra ? ntab_intra : 0;
    P.lbeta_inter = lbeta_inter;
    P.ntab_inter = lbeta_inter ? ntab_inter : 0;
    P.outl = outl;
    P.outl_thres = outl_thres;
    P.outl_stats = reinterpret_cast<unsigned long long *>(outl_stats);
    P.p = p;
    P.expcc = expcc;
    const long long ngroups = n >> 2;
    long long blocks = (ngroups + kPvalThreads - 1) / kPvalThreads;
    const long long cap = (long long)kNumSMs * 64;  // grid-stride beyond 64 CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (bias)
        pvalues_kernel<true><<<(unsigned int)blocks, kPvalThreads, 0, st>>>(P);
    else
        pvalues_kernel<false><<<(unsigned int)blocks, kPvalThreads, 0, st>>>(P);
    FHC_LAUNCH_CHECK("pvalues_kernel");
    return FHC_OK;
}

// Host builds of the table arithmetic (same source as the device code) so CPU-only tests can pin log_cr against
// libm's log and lbeta_cephes against scipy without a GPU.  Not used by the product path.
extern "C" double fhc_host_log_cr(double x) { return fhc::log_cr(x); }
extern "C" double fhc_host_lbeta(double a, double b) { return fhc::lbeta_cephes(a, b); }
