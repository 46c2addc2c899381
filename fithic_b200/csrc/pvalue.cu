// K3 -- fused per-contact significance: classify, bias gather, spline-table lookup, binomial survival, ExpCC.
//
// Replaces the per-line loop of fit_Spline (reference fithic/fithic.py:1017-1123) and scipy.special.bdtrc
// (:1070, :1101).  16 B/contact in (four 128-bit streaming loads per 4 contacts), 16 B/contact out (p and ExpCC as
// 128-bit stores).  The bias vector, the distance table and the lbeta table are small (<= 5 MB) and stay in L1/L2.
// The kernel is FP64-ALU bound (a continued fraction or a tail sum plus 4 transcendentals per contact), so what matters
// is keeping all 32 lanes of a warp on the same instruction.  A contact needs one of three very different evaluations
// (count == 1 closed form; tail sum when the count is below its expectation; continued fraction above it), each with a
// data-dependent trip count, which in a thread-per-contact kernel leaves ~8 of 32 lanes active (measured).  Instead a
// CTA works on a tile of 2048 contacts in phases:
//   1. load + classify + prior + ExpCC; cheap results go straight to shared memory, the rest is appended to one work
//      list per evaluation kind (warp-aggregated shared-memory appends);
//   2. continued fractions: each lane steps its own fraction; a lane that converges stores numerator/denominator and
//      takes the next contact from the list while its neighbours keep iterating (no lane waits for the slowest);
//   3. tail sums, same scheme;
//   4. one uniform pass turns every numerator/denominator into a p-value (3 logs + 1 exp, same code for both kinds),
//      another handles the closed forms;
//   5. p leaves through coalesced 128-bit stores; outliers are flagged.
#define FHC_PROFILE_STREAM st
#include <stdlib.h>

#include <string.h>

#include "pvalue_common.cuh"

namespace fhc {

constexpr int kPvalThreads = 256;
constexpr int kPvalTile = 2048;  // contacts per CTA tile (8 per thread)

struct PvalSmem {
    double x[kPvalTile];   // prior of the contacts that need real work
    double wp[kPvalTile];  // trivial p / value of the fraction or tail sum / finally p
    int cnt[kPvalTile];
    unsigned short work[kPvalTile];  // [0, nCf) continued fractions, [nCf, nCf + nTail) tail sums (local contact indices)
    unsigned short k0[kPvalTile];    // closed forms
    unsigned char inter[kPvalTile];  // 1: the contact is scored against N_inter
    unsigned long long warp_tot[kPvalThreads / 32 + 1];
    unsigned int cursor;             // next unclaimed entry of work[]
};

// claim the next entries of the work list for the lanes that need one (every lane of the warp must call);
// returns the position in the list or -1
__device__ __forceinline__ int work_claim(bool need, unsigned int total, unsigned int *cursor, int lane) {
    const unsigned int m = __ballot_sync(0xffffffffu, need);
    if (m == 0) return -1;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(cursor, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!need) return -1;
    const unsigned int k = base + __popc(m & ((1u << lane) - 1u));
    return k < total ? (int)k : -1;
}

constexpr unsigned long long kPack = 1ull << 21;  // three 21-bit counters in one 64-bit word

template <bool HAS_BIAS, int kMinCtas>
__global__ void __launch_bounds__(kPvalThreads, kMinCtas) pvalues_kernel(const PvalParams P) {
    extern __shared__ __align__(16) unsigned char pval_smem_raw[];
    PvalSmem &S = *reinterpret_cast<PvalSmem *>(pval_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ntiles = (P.n + kPvalTile - 1) / kPvalTile;
    unsigned int flagged = 0;
    const int *m1s = reinterpret_cast<const int *>(P.mid1), *m2s = reinterpret_cast<const int *>(P.mid2);
    const int *cs = reinterpret_cast<const int *>(P.cnt);
    const unsigned int *hs = reinterpret_cast<const unsigned int *>(P.chrs);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * kPvalTile;
        const bool full = base + kPvalTile <= P.n;
        if (tid == 0) S.cursor = 0;
        // ---- phase 1: load, classify, prior, ExpCC ----
        unsigned int codes = 0;  // 2 bits per contact of this thread: PvalClass
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int l0 = (h * kPvalThreads + tid) * 4;  // local index of this thread's 4 contacts
            int m1[4], m2[4], cc[4];
            unsigned int ch[4];
            if (full) {
                const long long g = (base + l0) >> 2;
                const int4 a1 = ldg_stream(P.mid1 + g), a2 = ldg_stream(P.mid2 + g), ac = ldg_stream(P.cnt + g);
                const int4 ah = ldg_stream(P.chrs + g);
                m1[0] = a1.x; m1[1] = a1.y; m1[2] = a1.z; m1[3] = a1.w;
                m2[0] = a2.x; m2[1] = a2.y; m2[2] = a2.z; m2[3] = a2.w;
                cc[0] = ac.x; cc[1] = ac.y; cc[2] = ac.z; cc[3] = ac.w;
                ch[0] = (unsigned int)ah.x; ch[1] = (unsigned int)ah.y; ch[2] = (unsigned int)ah.z; ch[3] = (unsigned int)ah.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long i = base + l0 + k;
                    const bool ok = i < P.n;
                    m1[k] = ok ? m1s[i] : 0;
                    m2[k] = ok ? m2s[i] : 0;
                    cc[k] = ok ? cs[i] : 0;
                    ch[k] = ok ? hs[i] : 0x00010000u;  // padding: an inter line
                }
            }
            double e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int li = l0 + k;
                double p, prior;
                bool use_inter;
                PvalClass cls = pval_prepare<HAS_BIAS>(P, m1[k], m2[k], cc[k], ch[k], p, e[k], prior, use_inter);
                if (!full && base + li >= P.n) cls = kClsDone;
                if (cls == kClsDone) {
                    S.wp[li] = p;
                } else {
                    S.x[li] = prior;
                    S.cnt[li] = cc[k];
                    S.inter[li] = use_inter ? 1 : 0;
                }
                codes |= (unsigned int)cls << (2 * (h * 4 + k));
            }
            if (full) {
                double2 *ee = reinterpret_cast<double2 *>(P.expcc + base + l0);
                __stcs(ee, make_double2(e[0], e[1]));
                __stcs(ee + 1, make_double2(e[2], e[3]));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (base + l0 + k < P.n) P.expcc[base + l0 + k] = e[k];
            }
        }
        // work lists by a CTA-wide exclusive scan of (cf, tail, k0) counts packed in one word: no atomics, and the
        // lists come out in contact order
        unsigned long long mine = 0;
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
            const unsigned int c = (codes >> (2 * s8)) & 3u;
            mine += (c == kClsCf ? 1ull : 0ull) + (c == kClsTail ? kPack : 0ull) + (c == kClsK0 ? kPack * kPack : 0ull);
        }
        unsigned long long inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) S.warp_tot[warp] = inc;
        __syncthreads();
        unsigned long long pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kPvalThreads / 32; ++w) {
            const unsigned long long t = S.warp_tot[w];
            if (w < warp) pre += t;
            tot += t;
        }
        const unsigned int nCf = (unsigned int)(tot & (kPack - 1)), nTail = (unsigned int)((tot >> 21) & (kPack - 1));
        const unsigned int nK0 = (unsigned int)(tot >> 42);
        {
            const unsigned long long ex = pre + inc - mine;
            unsigned int oCf = (unsigned int)(ex & (kPack - 1));
            unsigned int oTail = nCf + (unsigned int)((ex >> 21) & (kPack - 1));
            unsigned int oK0 = (unsigned int)(ex >> 42);
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) {
                const unsigned int c = (codes >> (2 * s8)) & 3u;
                const unsigned short li = (unsigned short)(((s8 >> 2) * kPvalThreads + tid) * 4 + (s8 & 3));
                if (c == kClsCf) S.work[oCf++] = li;
                if (c == kClsTail) S.work[oTail++] = li;
                if (c == kClsK0) S.k0[oK0++] = li;
            }
        }
        __syncthreads();
        // ---- phase 2: continued fractions and tail sums from one work list; a lane that finishes its contact takes the
        // next one while its neighbours keep iterating.  Only the warps that straddle nCf run both bodies.
        // (Tried and rejected on B200: preparing the per-contact set-up in a separate uniform pass, 16.4 ms instead of
        // 15.6 ms at 300 M contacts; restricting the loop to as many warps as stay full, 19.7 ms -- the kernel is bound by
        // the latency of the dependent FP64 chains, so more lanes in flight win over fuller warps.) ----
        const unsigned int nWork = nCf + nTail;
        {
            CfState st;
            int pos = -1;  // position in S.work
            int item = 0;
            bool exhausted = false;
            while (true) {
                const bool need = pos < 0 && !exhausted;
                const int got = work_claim(need, nWork, &S.cursor, lane);
                if (need) {
                    if (got < 0) {
                        exhausted = true;
                    } else {
                        pos = got;
                        item = S.work[pos];
                        const bool ui = S.inter[item] != 0;
                        const int N = ui ? P.N_inter : P.N_intra;
                        const double aa = (double)S.cnt[item], xx = S.x[item];
                        if ((unsigned int)pos < nCf) {
                            const double bb = (double)((long long)N - S.cnt[item] + 1);
                            cf_init(st, aa, bb, xx, cf_uses_d(aa, bb, xx));
                        } else {
                            double cN;
                            int M;
                            const double invN = ui ? P.invN_inter : P.invN_intra;
                            tail_prepare(aa, (double)N, invN, xx, __dsub_rn(1.0, xx), cN, M);
                            tail_load(st, aa, (double)N, invN, cN, M);
                        }
                    }
                }
                if (__ballot_sync(0xffffffffu, pos >= 0) == 0) break;
                bool done = false;
                if (pos >= 0) {
                    if ((unsigned int)pos < nCf)
                        done = cf_step(st);
                    else
                        done = tail_step(st);
                }
                if (done) {
                    S.wp[item] = st.pkm1 / st.qkm1;
                    pos = -1;
                }
            }
        }
        __syncthreads();
        // ---- phase 3: p-values (same code for both kinds: 3 log + 1 exp), closed forms ----
        for (unsigned int i = tid; i < nCf + nTail; i += kPvalThreads) {
            const bool tail = i >= nCf;
            const int item = S.work[i];
            const bool ui = S.inter[item] != 0;
            const int N = ui ? P.N_inter : P.N_intra;
            const int c = S.cnt[item];
            const double aa = (double)c, bb = (double)((long long)N - c + 1);
            const double *tab = ui ? P.lbeta_inter : P.lbeta_intra;
            const long long ntab = ui ? P.ntab_inter : P.ntab_intra;
            const double lb = (c < ntab) ? __ldg(tab + c) : lbeta_cephes(aa, bb);
            S.wp[item] = incbet_finish(tail, aa, bb, S.x[item], lb, S.wp[item]);
        }
        for (unsigned int i = tid; i < nK0; i += kPvalThreads) {
            const int item = S.k0[i];
            S.wp[item] = bdtrc_k0(S.inter[item] ? P.N_inter : P.N_intra, S.x[item]);
        }
        __syncthreads();
        // ---- phase 4: coalesced stores of p, outlier flags ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int l0 = (h * kPvalThreads + tid) * 4;
            const double p0 = S.wp[l0], p1 = S.wp[l0 + 1], p2 = S.wp[l0 + 2], p3 = S.wp[l0 + 3];
            if (full) {
                double2 *pp = reinterpret_cast<double2 *>(P.p + base + l0);
                __stcs(pp, make_double2(p0, p1));
                __stcs(pp + 1, make_double2(p2, p3));
                if (P.outl != nullptr) {
                    outlier_mark(P, base + l0 + 0, p0, flagged);
                    outlier_mark(P, base + l0 + 1, p1, flagged);
                    outlier_mark(P, base + l0 + 2, p2, flagged);
                    outlier_mark(P, base + l0 + 3, p3, flagged);
                }
            } else {
                const double pv[4] = {p0, p1, p2, p3};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long i = base + l0 + k;
                    if (i < P.n) {
                        P.p[i] = pv[k];
                        if (P.outl != nullptr) outlier_mark(P, i, pv[k], flagged);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (P.outl != nullptr) {
        const unsigned long long f = warp_sum((unsigned long long)flagged);
        if (lane == 0 && f) atomicAdd(P.outl_stats, f);
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

int pvalues_tile_launch(const PvalParams &P, cudaStream_t st) {
    const long long n = P.n;
    // resident CTAs per SM: 3 (80 registers, ~130 B of spills outside the loops; measured 15.6 ms at 300 M contacts) or
    // 2 (116 registers, no spills; 18.5 ms).  FHC_PVAL_OCC=2 selects the latter for experiments.
    static int occ = -1;
    if (occ < 0) {
        const char *e = getenv("FHC_PVAL_OCC");
        occ = (e && atoi(e) == 2) ? 2 : 3;
    }
    long long blocks = (n + kPvalTile - 1) / kPvalTile;
    const long long cap = (long long)kNumSMs * occ;  // persistent CTAs, grid-stride over tiles
    if (blocks > cap) blocks = cap;
    const size_t smem = sizeof(PvalSmem);
#define FHC_LAUNCH_PVAL(B, O)                                                                                           \
    do {                                                                                                                \
        FHC_CUDA(cudaFuncSetAttribute(pvalues_kernel<B, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        pvalues_kernel<B, O><<<(unsigned int)blocks, kPvalThreads, smem, st>>>(P);                                     \
    } while (0)
    if (P.bias) {
        if (occ == 3) FHC_LAUNCH_PVAL(true, 3); else FHC_LAUNCH_PVAL(true, 2);
    } else {
        if (occ == 3) FHC_LAUNCH_PVAL(false, 3); else FHC_LAUNCH_PVAL(false, 2);
    }
#undef FHC_LAUNCH_PVAL
    FHC_LAUNCH_CHECK("pvalues_kernel");
    return FHC_OK;
}

}  // namespace fhc

extern "C" int fhc_lbeta_table(int64_t N, double *tab, int64_t ntab, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(tab != nullptr && ntab > 0, FHC_E_INVALID, "fhc_lbeta_table: null table or ntab <= 0");
    FHC_REQUIRE(N >= 0 && N < (1ll << 31), FHC_E_RANGE,
                "fhc_lbeta_table: N = %lld does not fit the int32 that scipy.special.bdtrc truncates n to", (long long)N);
    const int threads = 128;
    const long long blocks = (ntab + threads - 1) / threads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    lbeta_table_kernel<<<(unsigned int)blocks, threads, 0, st>>>((int)N, tab, ntab);
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    bdtrc_kernel<<<(unsigned int)blocks, threads, 0, st>>>(
        cnt_minus_1, (int)N, prior, n, lbeta, lbeta ? ntab : 0, out);
    FHC_LAUNCH_CHECK("bdtrc_kernel");
    return FHC_OK;
}

// shared argument checks and parameter block of fhc_pvalues / fhc_pvalues_prepass
static int pval_params(fhc::PvalParams &P, const char *who, int32_t mode, const int32_t *mid1, const int32_t *mid2,
                       const int32_t *cnt, const uint32_t *chrs, const int64_t *run_start, const uint32_t *run_val, int32_t nruns,
                       int64_t n, const double *bias, const int32_t *bias_mid, const int64_t *chr_off, int32_t nchr,
                       int32_t bias_sparse, int32_t res, int64_t L, int64_t U, double tL, double tU, int64_t line_base,
                       bool need_contacts) {
    using namespace fhc;
    FHC_REQUIRE(mode == FHC_MODE_INTRA_ONLY || mode == FHC_MODE_INTER_ONLY || mode == FHC_MODE_ALL, FHC_E_INVALID,
                "%s: unknown mode %d", who, mode);
    FHC_REQUIRE(n >= 0 && res > 0, FHC_E_INVALID, "%s: need n >= 0 and res > 0", who);
    FHC_REQUIRE(n == 0 || !need_contacts || (mid1 && mid2), FHC_E_INVALID, "%s: null pointer", who);
    FHC_REQUIRE(n == 0 || !need_contacts || chrs != nullptr ||
                    (run_start && run_val && nruns >= 1 && nruns <= FHC_MAX_CHR_RUNS),
                FHC_E_INVALID, "%s: need chrs or 1 <= nruns <= %d chromosome runs", who, FHC_MAX_CHR_RUNS);
    FHC_REQUIRE(aligned16(mid1) && aligned16(mid2) && aligned16(cnt) && aligned16(chrs), FHC_E_INVALID,
                "%s: contact arrays must be 16-byte aligned", who);
    FHC_REQUIRE(bias == nullptr || (chr_off && nchr > 0), FHC_E_INVALID, "%s: bias needs chr_off and nchr", who);
    FHC_REQUIRE(!(bias && bias_sparse) || bias_mid != nullptr, FHC_E_INVALID,
                "%s: the sparse bias layout needs bias_mid (the sorted mid points)", who);
    FHC_REQUIRE(L >= -1 && U >= -1, FHC_E_INVALID, "%s: L and U must be >= -1", who);
    memset(&P, 0, sizeof(P));
    P.mode = mode;
    P.mid1 = reinterpret_cast<const int4 *>(mid1);
    P.mid2 = reinterpret_cast<const int4 *>(mid2);
    P.cnt = reinterpret_cast<const int4 *>(cnt);
    P.chrs = reinterpret_cast<const int4 *>(chrs);
    P.run_start = reinterpret_cast<const long long *>(run_start);
    P.run_val = run_val;
    P.nruns = chrs ? 0 : nruns;
    P.n = n;
    P.bias = bias;
    P.bias_mid = bias_mid;
    P.bias_sparse = bias_sparse ? 1 : 0;
    P.chr_off = reinterpret_cast<const long long *>(chr_off);
    P.nchr = nchr;
    P.res.d = (unsigned int)res;
    P.res.M = res == 1 ? 0ull : (~0ull) / (unsigned long long)res + 1ull;  // ceil(2^64 / res)
    P.Llo = L < 0 ? 0 : L;
    P.Uhi = U < 0 ? INT64_MAX : U;
    P.tL = tL;
    P.tU = tU;
    P.line_base = line_base;
    return FHC_OK;
}

extern "C" int fhc_pvalues_prepass(int32_t mode, const int32_t *mid1, const int32_t *mid2, const uint32_t *chrs,
                                   const int64_t *run_start, const uint32_t *run_val, int32_t nruns, int64_t n,
                                   const double *bias, const int32_t *bias_mid, const int64_t *chr_off, int32_t nchr,
                                   int32_t bias_sparse, int32_t res, int64_t L, int64_t U, double tL, double tU,
                                   int64_t line_base, uint32_t *code, double *b12, void *stream) {
    using namespace fhc;
    PvalParams P;
    const int rc = pval_params(P, "fhc_pvalues_prepass", mode, mid1, mid2, nullptr, chrs, run_start, run_val, nruns, n, bias,
                               bias_mid, chr_off, nchr, bias_sparse, res, L, U, tL, tU, line_base, true);
    if (rc != FHC_OK) return rc;
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(code && b12 && aligned16(code) && aligned16(b12), FHC_E_INVALID,
                "fhc_pvalues_prepass: code and b12 must be 16-byte aligned device arrays");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    return pvalues_prepass_launch(P, code, b12, st);
}

extern "C" int fhc_pvalues(int32_t mode, const int32_t *mid1, const int32_t *mid2, const int32_t *cnt,
                           const uint32_t *chrs, const int64_t *run_start, const uint32_t *run_val, int32_t nruns,
                           int64_t n, const double *bias, const int32_t *bias_mid,
                           const int64_t *chr_off, int32_t nchr, int32_t bias_sparse, int32_t res, int64_t L, int64_t U, const double *lut,
                           int64_t D, int64_t N_intra, int64_t N_inter, double interChrProb, double tL, double tU,
                           const double *lbeta_intra, int64_t ntab_intra, const double *lbeta_inter, int64_t ntab_inter,
                           uint8_t *outl, int64_t line_base, double outl_thres, uint64_t *outl_stats, double *p,
                           double *expcc, const uint32_t *pre_code, const double *pre_b12, void *workspace,
                           size_t workspace_bytes, void *stream) {
    using namespace fhc;
    PvalParams P;
    const bool pre = pre_code != nullptr;
    {
        const int rc = pval_params(P, "fhc_pvalues", mode, mid1, mid2, cnt, chrs, run_start, run_val, nruns, n, bias, bias_mid,
                                   chr_off, nchr, bias_sparse, res, L, U, tL, tU, line_base, !pre);
        if (rc != FHC_OK) return rc;
    }
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(cnt && p && expcc, FHC_E_INVALID, "fhc_pvalues: null pointer");
    FHC_REQUIRE(aligned16(p) && aligned16(expcc), FHC_E_INVALID, "fhc_pvalues: output arrays must be 16-byte aligned");
    FHC_REQUIRE((pre_code == nullptr) == (pre_b12 == nullptr) && aligned16(pre_code) && aligned16(pre_b12), FHC_E_INVALID,
                "fhc_pvalues: pre_code and pre_b12 come together (fhc_pvalues_prepass), 16-byte aligned");
    FHC_REQUIRE(!pre || D < (1ll << 30), FHC_E_INVALID, "fhc_pvalues: the pre-pass layout holds distance slots below 2^30");
    FHC_REQUIRE(mode == FHC_MODE_INTER_ONLY || (lut != nullptr && D > 0), FHC_E_INVALID,
                "fhc_pvalues: the distance table is required unless mode is interOnly");
    FHC_REQUIRE(N_intra >= 0 && N_intra < (1ll << 31) && N_inter >= 0 && N_inter < (1ll << 31), FHC_E_RANGE,
                "fhc_pvalues: N_intra = %lld / N_inter = %lld do not fit the int32 scipy.special.bdtrc truncates n to "
                "(the reference returns NaN or garbage there, SURVEY F5)",
                (long long)N_intra, (long long)N_inter);
    FHC_REQUIRE(outl == nullptr || outl_stats != nullptr, FHC_E_INVALID, "fhc_pvalues: outl needs outl_stats");
    P.pre_code = pre_code;
    P.pre_b12 = pre_b12;
    P.lut = lut;
    P.D = lut ? D : 0;
    P.N_intra = (int)N_intra;
    P.N_inter = (int)N_inter;
    P.invN_intra = N_intra > 0 ? 1.0 / (double)N_intra : 0.0;
    P.invN_inter = N_inter > 0 ? 1.0 / (double)N_inter : 0.0;
    P.interChrProb = interChrProb;
    P.lbeta_intra = lbeta_intra;
    P.ntab_intra = lbeta_intra ? ntab_intra : 0;
    P.lbeta_inter = lbeta_inter;
    P.ntab_inter = lbeta_inter ? ntab_inter : 0;
    P.outl = outl;
    P.outl_thres = outl_thres;
    P.outl_stats = reinterpret_cast<unsigned long long *>(outl_stats);
    P.p = p;
    P.expcc = expcc;
    // Two implementations (same arithmetic, parity-tested against each other and the oracle):
    //   work lists (pvalue_lists.cu): needs the caller's workspace; default whenever one is passed
    //   tile-phased single kernel (below): no workspace; FHC_PVAL_IMPL=tile forces it for experiments
    const char *impl_env = getenv("FHC_PVAL_IMPL");  // read per call so a test can run both in one process
    const int impl = (impl_env && impl_env[0] == 't') ? 1 : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    if (workspace != nullptr && impl == 0) return pvalues_lists_launch(P, workspace, workspace_bytes, st);
    FHC_REQUIRE(!pre, FHC_E_INVALID, "fhc_pvalues: the tile-phased kernel does not read a pre-pass (pass a workspace)");
    FHC_REQUIRE(chrs != nullptr && mid1 != nullptr && mid2 != nullptr, FHC_E_INVALID,
                "fhc_pvalues: the tile-phased kernel needs the mid1 / mid2 / chrs arrays (chromosome runs are read by the "
                "work-list pipeline: pass a workspace)");
    return pvalues_tile_launch(P, st);
}

extern "C" size_t fhc_pvalues_workspace_bytes(int64_t n, int64_t ntab) {
    return fhc::pvalues_lists_workspace_bytes(n < 0 ? 0 : n, ntab < 0 ? 0 : ntab);
}

// Host builds of the table arithmetic (same source as the device code) so CPU-only tests can pin log_cr against
// libm's log and lbeta_cephes against scipy without a GPU.  Not used by the product path.
extern "C" double fhc_host_log_cr(double x) { return fhc::log_cr(x); }
extern "C" double fhc_host_lbeta(double a, double b) { return fhc::lbeta_cephes(a, b); }
// scipy.special.bdtrc(count - 1, N, prior) through the arithmetic of the work-list pipeline / of the tile kernel
extern "C" double fhc_host_bdtrc_lists(int32_t count, int64_t N, double prior) {
    return fhc::bdtrc_lists_scalar(count, (int)N, prior);
}
extern "C" double fhc_host_one_minus_exp(double y) { return fhc::one_minus_exp(y); }
// The lower tail sum of the swapped incbet branch for one contact, as numerator and denominator: in_place != 0 as
// pval_finish_kernel sums the short ones (tail_short_sum), else as the queue of pval_iterate_kernel does (tail_fwd_*).
extern "C" void fhc_host_tail_sum(int32_t count, int64_t N, double prior, int32_t in_place, double *num, double *den) {
    const double dN = (double)N;
    const double cN = fhc::tail_cn(dN, prior, fhc::rn_sub(1.0, prior));
    if (in_place) {
        *num = fhc::tail_short_sum((double)count, dN, 1.0 / dN, cN, den);
        return;
    }
    fhc::CfState s;
    fhc::tail_fwd_load(s, (double)count, dN, 1.0 / dN, cN);
    while (!fhc::tail_fwd_step(s)) {
    }
    *num = s.pkm1;
    *den = s.qkm1;
}
