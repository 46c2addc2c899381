// Device arithmetic for scipy.special.bdtrc as Fit-Hi-C calls it (reference fithic/fithic.py:1070, :1101).
//
// scipy's implementation is xsf::cephes::{bdtrc, incbet, incbcf, incbd, incbet_pseries, lbeta, lgam} (third party,
// not in /root/reference).  What has to be REPRODUCED and what is free to be re-derived:
//   * lbeta(a, b) with a + b = N + 1 cancels two ~N log N terms (lgam(N - c + 1) - lgam(N + 1)); at N ~ 1e9 its
//     rounding noise is ~4e-6 relative in the p-value.  Parity at 1e-6 therefore needs cephes' formula
//     with the same roundings: every operation below that feeds it is an explicit round-to-nearest intrinsic (no FMA
//     contraction) and log() is evaluated in double-double and rounded once, which is what glibc's log returns
//     (checked on 20,000 integers in [1e8, 2^31]: identical).  lbeta depends only on (count, N): it is tabulated once
//     per run by lbeta_table_kernel and gathered per contact.
//   * 1 - x is rounded to double before log() in cephes (b*log(1-x) with b ~ 1e9 amplifies that rounding to ~1e-7);
//     the same subtraction is done here.
//   * the continued fractions are evaluated with an equivalence transform (no division inside the loop), stop at a
//     relative change of kCfTol instead of cephes' 3 ulp / 300 iterations (cephes oscillates at rounding level: 46% of
//     real inputs hit the 300 cap), and cover the power-series region too.  All of that moves the result
//     by <= ~1e-10 relative, far inside the 1e-6 contract.
#pragma once
#include <math.h>
#include <string.h>

#include "common.cuh"

#if defined(__CUDA_ARCH__)
#define FHC_HD __host__ __device__ __forceinline__
#else
#define FHC_HD __host__ __device__ inline
#endif

namespace fhc {

constexpr double kMACHEP = 1.11022302462515654042E-16;
constexpr double kMAXLOG = 7.09782712893383996843E2;
constexpr double kMINLOG = -7.451332191019412076235E2;
constexpr double kMAXGAM = 171.624376956302725;
constexpr double kLS2PI = 0.91893853320467274178;
constexpr double kASYMP = 1e6;
constexpr double kCfTol = 1e-11;
constexpr int kCfMaxIter = 300;

// ---- exactly rounded primitives (identical on host and device) ----------------------------------------------------
FHC_HD double rn_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;  // host objects are built with -ffp-contract=off
#endif
}
FHC_HD double rn_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
FHC_HD double rn_sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
FHC_HD double rn_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
FHC_HD double rn_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// bit access: device intrinsics, memcpy on the host (the host build exists for CPU-only tests of the same source)
FHC_HD int dbl_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    unsigned long long b;
    memcpy(&b, &x, sizeof(b));
    return (int)(b >> 32);
#endif
}
FHC_HD int dbl_lo(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    unsigned long long b;
    memcpy(&b, &x, sizeof(b));
    return (int)(b & 0xffffffffull);
#endif
}
FHC_HD double dbl_from_hi(int hi) {  // low word 0
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, 0);
#else
    const unsigned long long b = (unsigned long long)(unsigned int)hi << 32;
    double x;
    memcpy(&x, &b, sizeof(x));
    return x;
#endif
}

// ---- double-double arithmetic (only used while tabulating lbeta) ----------------------------------------------------
struct dd {
    double hi, lo;
};
FHC_HD dd two_sum(double a, double b) {
    const double s = rn_add(a, b);
    const double bb = rn_sub(s, a);
    const double e = rn_add(rn_sub(a, rn_sub(s, bb)), rn_sub(b, bb));
    return {s, e};
}
FHC_HD dd quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = rn_add(a, b);
    return {s, rn_sub(b, rn_sub(s, a))};
}
FHC_HD dd two_prod(double a, double b) {
    const double p = rn_mul(a, b);
    return {p, rn_fma(a, b, -p)};
}
FHC_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = rn_add(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = rn_add(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
FHC_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = rn_add(p.lo, rn_add(rn_mul(a.hi, b.lo), rn_mul(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}
FHC_HD dd dd_div(dd a, dd b) {
    const double q1 = rn_div(a.hi, b.hi);
    const dd t1 = dd_mul(b, dd{q1, 0.0});
    dd r = dd_add(a, dd{-t1.hi, -t1.lo});
    const double q2 = rn_div(r.hi, b.hi);
    const dd t = dd_mul(b, dd{q2, 0.0});
    r = dd_add(r, dd{-t.hi, -t.lo});
    const double q3 = rn_div(r.hi, b.hi);
    dd q = quick_two_sum(q1, q2);
    return dd_add(q, dd{q3, 0.0});
}

// log(x), x > 0 finite, evaluated to ~2^-100 and rounded once to double.
// x = m 2^e with m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716, 26 odd terms.
#if defined(__CUDA_ARCH__)
__host__ __device__ __noinline__
#else
inline
#endif
double log_cr(double x) {
    int e;
    double m = frexp(x, &e);  // m in [0.5, 1)
    if (m < 0.70710678118654752440) {
        m = rn_mul(m, 2.0);
        e -= 1;
    }
    const dd num = {rn_sub(m, 1.0), 0.0};  // exact for m in [0.5, 2]
    const dd den = two_sum(m, 1.0);
    const dd s = dd_div(num, den);
    const dd s2 = dd_mul(s, s);
    dd acc = dd_div(dd{1.0, 0.0}, dd{53.0, 0.0});
    for (int k = 25; k >= 0; --k) {
        const dd ck = dd_div(dd{1.0, 0.0}, dd{(double)(2 * k + 1), 0.0});
        acc = dd_add(dd_mul(acc, s2), ck);
    }
    dd r = dd_mul(s, acc);
    r = dd{rn_mul(r.hi, 2.0), rn_mul(r.lo, 2.0)};
    // e * ln2 with ln2 = hi + lo + lo2 (three-double constant)
    const double ed = (double)e;
    dd el = two_prod(ed, 0.69314718055994528623);             // 0x3FE62E42FEFA39EF
    el.lo = rn_add(el.lo, rn_mul(ed, 2.3190468138462995584e-17));  // 0x3C7ABC9E3B39803F
    el = quick_two_sum(el.hi, el.lo);
    el = dd_add(el, dd{rn_mul(ed, 5.7077084384162120658e-34), 0.0});  // third limb of ln2
    r = dd_add(el, r);
    return rn_add(r.hi, r.lo);
}

#define FHC_LOG_FN log_cr
#include "cephes_lbeta.inc"
#undef FHC_LOG_FN

#if defined(__CUDACC__)
// ---- continued fractions of the incomplete beta function, division free ---------------------------------------------
// cephes incbcf (use_d = false, z = x) and incbd (use_d = true, z = x / (1 - x)):
//   p_j = p_{j-1} + (n_j / d_j) p_{j-2}  is carried as  P_j = d_j P_{j-1} + d_{j-1} n_j P_{j-2}  (same for Q), so
// P_j / Q_j is unchanged and no division is needed until the end; convergence is tested by cross multiplication.
// The state is explicit so that a warp can keep all 32 lanes busy: a lane whose fraction has converged picks up the
// next contact from a shared work list while its neighbours keep iterating (pvalue.cu).
struct CfState {
    double a, apb, bm1, z;  // parameters: a, a + b, b - 1, x or x / (1 - x)
    double fi;              // iteration counter as a double
    double pkm2, qkm2, pkm1, qkm1, dprev;  // value = pkm1 / qkm1
    bool use_d;
};

FHC_HD void cf_init(CfState &s, double a, double b, double x, bool use_d) {
    s.a = a; s.apb = a + b; s.bm1 = b - 1.0; s.use_d = use_d;
    s.z = use_d ? x / (1.0 - x) : x;
    s.fi = 0.0;
    s.pkm2 = 0.0; s.qkm2 = 1.0; s.pkm1 = 1.0; s.qkm1 = 1.0; s.dprev = 1.0;
}
// z by a reciprocal instead of a full division (the work-list pipeline prepares it for a whole chunk at once)
FHC_HD double cf_z(double x, bool use_d) { return use_d ? x * (1.0 / (1.0 - x)) : x; }
// the same from a precomputed z
FHC_HD void cf_load(CfState &s, double a, double b, double z, bool use_d) {
    s.a = a; s.apb = a + b; s.bm1 = b - 1.0; s.use_d = use_d;
    s.z = z;
    s.fi = 0.0;
    s.pkm2 = 0.0; s.qkm2 = 1.0; s.pkm1 = 1.0; s.qkm1 = 1.0; s.dprev = 1.0;
}

// one iteration (two recurrence steps, updated in place: after step 1 the slot "km2" holds the newest convergent, after
// step 2 "km1" does again); returns true when the fraction has converged (or hit the iteration cap)
FHC_HD bool cf_step(CfState &s) {
    const double k1 = s.a + s.fi, k3 = k1 + s.fi, k4 = k3 + 1.0, k5 = 1.0 + s.fi, k8 = k3 + 2.0;
    const double up = s.apb + s.fi, dn = s.bm1 - s.fi;
    const double k2 = s.use_d ? dn : up, k6 = s.use_d ? up : dn;
    const double pprev = s.pkm1, qprev = s.qkm1;  // previous convergent (cephes' `ans`)
    const double d1 = k3 * k4;
    const double a1 = -(s.z * k1 * k2) * s.dprev;
    s.pkm2 = d1 * s.pkm1 + a1 * s.pkm2;
    s.qkm2 = d1 * s.qkm1 + a1 * s.qkm2;
    const double d2 = k4 * k8;
    const double a2 = (s.z * k5 * k6) * d1;
    s.pkm1 = d2 * s.pkm2 + a2 * s.pkm1;
    s.qkm1 = d2 * s.qkm2 + a2 * s.qkm1;
    s.dprev = d2;
    s.fi += 1.0;
    // |pprev/qprev - pk/qk| < tol |pk/qk|, by cross multiplication
    const double lhs = fabs(pprev * s.qkm1 - s.pkm1 * qprev), rhs = kCfTol * fabs(qprev * s.pkm1);
    if (lhs < rhs || s.fi >= (double)kCfMaxIter) return true;
    const double mag = fabs(s.qkm1) + fabs(s.pkm1);
    if (mag > 1.2676506002282294e30 || mag < 7.8886090522101181e-31) {  // 2^100, 2^-100: renormalise exactly
        const int ex = ((dbl_hi(mag) >> 20) & 0x7ff);
        if (ex != 0 && ex != 0x7ff) {
            const double sc = dbl_from_hi((2046 - ex) << 20);  // 2^-(ex-1023)
            s.pkm2 *= sc; s.pkm1 *= sc; s.qkm2 *= sc; s.qkm1 *= sc;
        }
    }
    return false;
}

// Lower binomial tail relative to its last term, for the swapped branch of incbet (x > a/(a+b): the observed count is
// below its expectation).  cephes evaluates I_{1-x}(b, a) there with the same continued fraction, but with a ~ N ~ 1e9
// that fraction no longer reaches 3 ulp: 94-100 % of such inputs run into the 300-iteration cap and land within ~1e-7
// of the true value (measured against Boost).  The tail is a short, all-positive finite sum instead:
//   P(X <= k) = pmf(k) (1 + s_k (1 + s_{k-1} (1 + ...))),   s_j = pmf(j-1)/pmf(j) = j (1-x) / ((N-j+1) x),  k = count-1
// evaluated inside-out as a ratio P/Q (no division in the loop).  Terms below 1e-17 of the sum are skipped: the term i
// steps below k is <= r^i exp(-i(i-1)/2k) with r = k/(N x) <= 1, which bounds the number of terms M.
// The sum S = P/Q gives  I_{1-x}(b, a) = (1-x)^b x^a / (b B(a, b)) * (S / x),  S / x being what cephes' fraction stands for.
// The state lives in the fields of a CfState (a lane works on one kind at a time, pvalue.cu):
//   P = pkm1, Q = qkm1 (so the value is pkm1 / qkm1 for both kinds), j = fi, d = dprev, cN = z, invN = a, terms left = bm1
// number of terms and the per-term factor (1-x)/(x N): the expensive part of the set-up
FHC_HD void tail_prepare(double count, double N, double invN, double x, double one_minus_x, double &cN,
                                             int &M) {
    const double k = count - 1.0;
    cN = (one_minus_x / x) * invN;
    double Md = k;
    if (k > 24.0) {  // for short sums the bound cannot beat k by much: skip the log and the square root
        const double r = k / (N * x);
        if (r > 0.0 && r < 1.0) Md = fmin(Md, ceil(-39.2 / log(r)) + 1.0);
        Md = fmin(Md, ceil(sqrt(78.4 * k)) + 1.0);
    }
    M = (int)Md;
}
FHC_HD void tail_load(CfState &s, double count, double N, double invN, double cN, int M) {
    s.a = invN;
    s.z = cN;
    s.pkm1 = 1.0; s.qkm1 = 1.0;
    s.fi = count - (double)M;  // k - M + 1
    s.dprev = (N - s.fi + 1.0) * invN;
    s.bm1 = (double)M;
}
FHC_HD void tail_init(CfState &s, double count, double N, double x, double one_minus_x) {
    double cN;
    int M;
    const double invN = 1.0 / N;
    tail_prepare(count, N, invN, x, one_minus_x, cN, M);
    tail_load(s, count, N, invN, cN, M);
}

// one term; returns true when the sum is complete
FHC_HD bool tail_step(CfState &s) {
    if (s.bm1 <= 0.0) return true;
    const double n = s.fi * s.z;
    const double dq = s.dprev * s.qkm1;
    s.pkm1 = fma(n, s.pkm1, dq);
    s.qkm1 = dq;
    s.fi += 1.0;
    s.dprev -= s.a;
    s.bm1 -= 1.0;
    if (s.qkm1 < 7.8886090522101181e-31) {  // 2^-100: only reachable when count is a sizeable fraction of N
        s.pkm1 *= 1.2676506002282294e30;
        s.qkm1 *= 1.2676506002282294e30;
    }
    return s.bm1 <= 0.0;
}

// which evaluation a contact needs once the cheap exits of bdtrc / incbet are taken
enum PvalClass : int { kClsDone = 0, kClsK0 = 1, kClsTail = 2, kClsCf = 3 };

// The branch structure of scipy.special.bdtrc(k = count - 1, n = N, p = prior) and of cephes incbet up to the point
// where real work starts.  Returns the class; for kClsDone `value` is the result.
FHC_HD PvalClass bdtrc_classify(int count, int N, double prior, double &value) {
    value = NAN;
    if (isnan(prior)) return kClsDone;
    if (prior < 0.0 || prior > 1.0) return kClsDone;
    const long long k = (long long)count - 1;
    if (k < 0) { value = 1.0; return kClsDone; }
    if ((long long)N < k) return kClsDone;
    if (k == (long long)N) { value = 0.0; return kClsDone; }
    if (k == 0) return kClsK0;
    if (prior <= 0.0) { value = 0.0; return kClsDone; }  // incbet: xx == 0
    if (prior >= 1.0) { value = 1.0; return kClsDone; }  // incbet: xx == 1
    const double aa = (double)count, bb = (double)((long long)N - k);
    return prior > aa / (aa + bb) ? kClsTail : kClsCf;
}

// k == 0: 1 - (1 - p)^n in cephes' two forms
FHC_HD double bdtrc_k0(int N, double prior) {
    const double dn = (double)N;
    if (prior < 0.01) {
        // -expm1(y): for y <= -1 there is no cancellation in 1 - exp(y), and that form rounds like libm's expm1 where
        // CUDA's expm1 saturates to -1 one ulp early (y in (-37.4, -36.7) decides p == 1.0 versus 1 - 2^-53)
        const double y = dn * log1p(-prior);
        return y <= -1.0 ? 1.0 - exp(y) : -expm1(y);
    }
    return 1.0 - pow(1.0 - prior, dn);
}

FHC_HD bool cf_uses_d(double aa, double bb, double xx) {
    return !(xx * (aa + bb - 2.0) - (aa - 1.0) < 0.0);  // cephes: y < 0 -> incbcf, else incbd
}

// last step of incbet: prefactor in log space times the fraction / tail value w
FHC_HD double incbet_finish(bool tail, double aa, double bb, double xx, double lbeta_ab, double w) {
    const double w1 = rn_sub(1.0, xx);  // cephes rounds 1 - x before taking its log
    double div;
    if (tail) {
        w = w / xx;
        div = bb;
    } else {
        if (cf_uses_d(aa, bb, xx)) w = w / w1;
        div = aa;
    }
    // log(w1): w1 - 1 is exact, and for the tiny priors of sparse maps a four-term series of log1p is exact to 1e-21
    const double y1 = w1 - 1.0;
    const double logw1 = (y1 > -7.62939453125e-06)
                             ? y1 * (1.0 + y1 * (-0.5 + y1 * (0.33333333333333331 + y1 * -0.25)))
                             : log(w1);
    double t = aa * log(xx) + bb * logw1 - lbeta_ab + log(w / div);
    t = t < kMINLOG ? 0.0 : exp(t);
    if (tail) t = (t <= kMACHEP) ? 1.0 - kMACHEP : 1.0 - t;
    return t;
}

// The same lower tail summed FORWARD from its largest term, for the work-list pipeline: no term count is needed up front
// (tail_prepare's log and square root ran on one lane in thirty), the sum simply stops when a term no longer matters.
//   T_0 = 1, T_i = T_{i-1} s_{k-i+1}, S = sum T_i;   s_j = n_j / d_j, n_j = j cN, d_j = (N-j+1)/N, cN = (1-x)/(x N)
// carried division free as A (numerator of the current term), Q (common denominator), P (numerator of the sum):
//   A <- A n_j,  Q <- Q d_j,  P <- P d_j + A,   S = P / Q.
// n_j < 1 throughout (the count is below its expectation in this branch) and d_j ~ 1, so nothing grows.  The terms fall at
// least like r (r - 1/lambda) (r - 2/lambda) ..., r = k / lambda < 1: what is left after a term A is below ~1.3 sqrt(lambda)
// A, so stopping at A < 1e-21 P leaves a relative error below 1e-17 for every lambda < 2^31.
// State in a CfState: P = pkm1, Q = qkm1 (value = pkm1 / qkm1 as for the fractions), A = pkm2, j = fi, d_j = dprev, cN = z,
// 1/N = a.
FHC_HD double tail_cn(double N, double x, double one_minus_x) { return one_minus_x * (1.0 / (x * N)); }
// Sums of at most kTailInline terms (counts up to kTailInline + 1) stay out of the work queue of the iterate kernel, where
// handing an item to a lane costs ~100 instructions (and only ~7 of 32 lanes take one per trip) against ~10 per term:
// pval_finish_kernel runs the same recurrence in place for them.  A warp pays for the longest sum among its lanes -- on the
// bench input 1.8 terms per item on average, ~10 trips per warp with a limit of 16 -- which is still a seventh of what the
// queue costs per item.  Measured (B200, 300 M contacts; iterate + finish): everything queued 1.83 + 1.25 ms; limit 4
// 1.49 + 1.50; limit 16 with the queue's own step function (renormalisation included) 0.81 + 1.96.
constexpr int kTailInline = 32;
FHC_HD bool tail_is_short(int count) { return count - 1 <= kTailInline; }
// tail_fwd_load + tail_fwd_step (below) for such a sum: the same operations on the same operands, without the
// renormalisation, which cannot trigger here (Q is a product of at most kTailInline factors (N - j + 1) / N with
// j <= kTailInline < N: never below 33! / 33^32 = 2e-12, far from 2^-100).  Returns P; Q in *q.
FHC_HD double tail_short_sum(double count, double N, double invN, double cN, double *q) {
    double A = 1.0, P = 1.0, Q = 1.0;
    double fi = count - 1.0;
    double d = (N - fi + 1.0) * invN;
    bool live = fi > 0.0;
    while (live) {
        A *= fi * cN;
        Q *= d;
        P = fma(P, d, A);
        fi -= 1.0;
        d += invN;
        live = !(fi <= 0.0 || A < 1e-21 * P);
    }
    *q = Q;
    return P;
}
FHC_HD void tail_fwd_load(CfState &s, double count, double N, double invN, double cN) {
    s.a = invN;
    s.z = cN;
    s.pkm1 = 1.0; s.qkm1 = 1.0; s.pkm2 = 1.0;
    s.fi = count - 1.0;                        // k: the first ratio is s_k
    s.dprev = (N - s.fi + 1.0) * invN;         // d_k
}
FHC_HD bool tail_fwd_step(CfState &s) {  // one term; returns true when the sum is complete
    if (s.fi <= 0.0) return true;
    s.pkm2 *= s.fi * s.z;
    s.qkm1 *= s.dprev;
    s.pkm1 = fma(s.pkm1, s.dprev, s.pkm2);
    s.fi -= 1.0;
    s.dprev += s.a;
    if (s.qkm1 < 7.8886090522101181e-31) {  // 2^-100: only reachable when the count is a sizeable fraction of N
        s.pkm2 *= 1.2676506002282294e30;
        s.pkm1 *= 1.2676506002282294e30;
        s.qkm1 *= 1.2676506002282294e30;
    }
    return s.fi <= 0.0 || s.pkm2 < 1e-21 * s.pkm1;
}

// ---- the forms used by the work-list pipeline (pvalue_lists.cu) -------------------------------------------------------
// Taylor coefficients of (expm1(r) - r) / r^2, highest degree first (1/14! ... 1/2!).  On the device they sit in constant
// memory so that each DFMA takes its coefficient as an operand instead of two moves of a 64-bit immediate.
#define FHC_EXPM1_TAYLOR                                                                                        \
    {1.1470745597729725e-11, 1.6059043836821613e-10, 2.0876756987868100e-09, 2.5052108385441720e-08,            \
     2.7557319223985893e-07, 2.7557319223985888e-06, 2.4801587301587302e-05, 1.9841269841269841e-04,            \
     1.3888888888888889e-03, 8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5}
static __constant__ double kExpm1TaylorDev[13] = FHC_EXPM1_TAYLOR;
static const double kExpm1TaylorHost[13] = FHC_EXPM1_TAYLOR;
FHC_HD double expm1_taylor(int i) {
#if defined(__CUDA_ARCH__)
    return kExpm1TaylorDev[i];
#else
    return kExpm1TaylorHost[i];
#endif
}

// 1 - exp(y) for y <= 0, one code path for every y: exp(y) = 2^k (1 + P(r)), r = y - k ln2, |r| <= ln2/2, P = expm1(r) by
// its Taylor polynomial (degree 14: 4e-18 relative).  k == 0 returns -P (full relative accuracy for tiny y); otherwise
// 1 - 2^k (1 + P) in one fused rounding.  The result is 1.0 from y ~ -37.4 on, like 1 - exp(y) in libm arithmetic.
FHC_HD double one_minus_exp(double y) {
    y = y < -100.0 ? -100.0 : y;  // (y is never NaN here: the prior was checked)
    const double t = fma(y, 1.4426950408889634074, 6755399441055744.0);  // 1.5 * 2^52: round to nearest integer
    const int k = dbl_lo(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, y);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double q = expm1_taylor(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 13; ++i) q = fma(q, r, expm1_taylor(i));
    const double P = fma(q * r, r, r);
    const double s = dbl_from_hi((1023 + k) << 20);  // 2^k, k in [-145, 0]
    return k == 0 ? 0.0 - P : fma(-s, 1.0 + P, 1.0);  // 0.0 - P: +0.0 for y == -0.0, like -expm1(-0.0)
}

// scipy.special.bdtrc(0, N, prior) = 1 - (1 - prior)^N, cephes' form for p < .01 (bdtr.h: -expm1(n log1p(-p))) with
// log1p by four terms of its series: exact to 7e-22 below 2^-17.  Larger priors take bdtrc_k0.
FHC_HD bool k0_series_ok(double prior) { return prior < 7.62939453125e-06; }
FHC_HD double bdtrc_k0_series(double dn, double prior) {
    const double l = -prior * fma(prior, fma(prior, fma(prior, 0.25, 0.33333333333333331), 0.5), 1.0);
    return one_minus_exp(dn * l);
}

// The prefactor of incbet in log space (cephes incbet.h) with the divisions folded into the exponent:
//   continued fraction  x^a (1-x)^b / (a B(a,b)) * w [/ (1-x) for incbd]   -> sub_cf = lbeta + log a
//   tail sum            x^a (1-x)^b / (b B(a,b)) * w / x                    -> sub_tail = lbeta + log b
FHC_HD double incbet_finish_folded(bool tail, double aa, double bb, double xx, double sub_cf, double sub_tail, double w) {
    const double w1 = rn_sub(1.0, xx);  // cephes rounds 1 - x before taking its log
    const double y1 = w1 - 1.0;         // exact
    const double logw1 = (y1 > -7.62939453125e-06)
                             ? y1 * (1.0 + y1 * (-0.5 + y1 * (0.33333333333333331 + y1 * -0.25)))
                             : log(w1);
    double ea = aa, eb = bb, sub = sub_cf;
    if (tail) {
        ea = aa - 1.0;
        sub = sub_tail;
    } else if (cf_uses_d(aa, bb, xx)) {
        eb = bb - 1.0;
    }
    double t = ea * log(xx) + eb * logw1 - sub + log(w);
    t = t < kMINLOG ? 0.0 : exp(t);
    if (tail) t = (t <= kMACHEP) ? 1.0 - kMACHEP : 1.0 - t;
    return t;
}

// Host + device scalar evaluation of the work-list pipeline's arithmetic (classification as in front_prepare, iteration,
// folded finish): what pval_front / pval_iterate / pval_finish compute for one contact.  Used by the CPU tests through
// fhc_host_bdtrc_lists and by nothing on the product path.
FHC_HD double bdtrc_lists_scalar(int count, int N, double prior) {
    double v;
    const PvalClass cls0 = bdtrc_classify(count, N, prior, v);
    if (cls0 == kClsDone) return v;
    if (cls0 == kClsK0) return k0_series_ok(prior) ? bdtrc_k0_series((double)N, prior) : bdtrc_k0(N, prior);
    const double aa = (double)count, bb = (double)((long long)N - count + 1);
    const bool tail = rn_mul(prior, (double)N + 1.0) > aa;  // front_prepare's division-free form of x > a / (a + b)
    const double lb = lbeta_cephes(aa, bb);
    CfState s;
    if (tail && tail_is_short(count)) {  // as pval_finish_kernel sums it in place
        double q;
        s.pkm1 = tail_short_sum(aa, (double)N, 1.0 / (double)N, tail_cn((double)N, prior, rn_sub(1.0, prior)), &q);
        s.qkm1 = q;
    } else if (tail) {
        tail_fwd_load(s, aa, (double)N, 1.0 / (double)N, tail_cn((double)N, prior, rn_sub(1.0, prior)));
        while (!tail_fwd_step(s)) {
        }
    } else {
        const bool use_d = cf_uses_d(aa, bb, prior);
        cf_load(s, aa, bb, cf_z(prior, use_d), use_d);
        while (!cf_step(s)) {
        }
    }
    return incbet_finish_folded(tail, aa, bb, prior, lb + log(aa), lb + log(bb), s.pkm1 / s.qkm1);
}

// scalar evaluation (one contact start to finish): used by the element-wise test entry point
__device__ __forceinline__ double bdtrc_dev(int count, int N, double prior, const double *__restrict__ lbeta_tab,
                                            long long ntab) {
    double v;
    const PvalClass cls = bdtrc_classify(count, N, prior, v);
    if (cls == kClsDone) return v;
    if (cls == kClsK0) return bdtrc_k0(N, prior);
    const double aa = (double)count, bb = (double)((long long)N - count + 1);
    const double lb = (count < ntab) ? __ldg(lbeta_tab + count) : lbeta_cephes(aa, bb);
    CfState s;
    if (cls == kClsTail) {
        tail_init(s, aa, (double)N, prior, rn_sub(1.0, prior));
        while (!tail_step(s)) {
        }
        return incbet_finish(true, aa, bb, prior, lb, s.pkm1 / s.qkm1);
    }
    cf_init(s, aa, bb, prior, cf_uses_d(aa, bb, prior));
    while (!cf_step(s)) {
    }
    return incbet_finish(false, aa, bb, prior, lb, s.pkm1 / s.qkm1);
}
#endif  // __CUDACC__

}  // namespace fhc
