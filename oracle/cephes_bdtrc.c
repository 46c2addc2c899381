/*
 * ORACLE (test infrastructure, NOT product code) -- CPU restatement of scipy.special.bdtrc.
 *
 * The reference calls scipy.special.bdtrc at fithic/fithic.py:1070 and :1101.  That arithmetic is not in
 * /root/reference: it is the third-party dependency scipy (unpinned in the reference's setup.py:25; scipy 1.18.1 is
 * installed in this image), whose implementation is xsf::cephes::{bdtrc, incbet, incbet_pseries, incbcf, incbd,
 * lbeta, lgam_sgn, Gamma} inside scipy/special/_ufuncs*.so.  This file restates the published Cephes algorithm
 * (Moshier, cephes/cprob/{bdtr,incbet}.c, cephes/cprob/gamma.c, scipy's beta.h additions) step for step, with the
 * reference's FULL stopping rule (3*MACHEP or 300 iterations) -- it is the checker, not the thing measured.
 *
 * Pinned by tests/test_oracle_cephes.py against golden vectors generated from scipy.special.bdtrc itself
 * (tests/golden/make_golden.py) and, where scipy is importable, against scipy live.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off cephes_bdtrc.c -o liboracle_cephes.so -lm
 * (-ffp-contract=off: scipy's x86-64 wheels carry no FMA contraction; keep the same rounding.)
 */
#include <math.h>
#include <stdint.h>

static const double MACHEP = 1.11022302462515654042E-16;
static const double MAXLOG = 7.09782712893383996843E2;
static const double MINLOG = -7.451332191019412076235E2;
static const double MAXGAM = 171.624376956302725;
static const double BIG = 4.503599627370496e15;
static const double BIGINV = 2.22044604925031308085e-16;
static const double LS2PI = 0.91893853320467274178;
static const double MAXLGM = 2.556348e305;
static const double ASYMP_FACTOR = 1e6;
static const double MAXSTIR = 143.01608;
static const double SQTPI = 2.50662827463100050242E0;

static double polevl(double x, const double *c, int n) {
    double a = c[0];
    for (int i = 1; i <= n; i++) a = a * x + c[i];
    return a;
}
static double p1evl(double x, const double *c, int n) {
    double a = x + c[0];
    for (int i = 1; i < n; i++) a = a * x + c[i];
    return a;
}

/* ---- Gamma / lgam (cephes gamma.c) ---- */
static const double GP[] = {1.60119522476751861407E-4, 1.19135147006586384913E-3, 1.04213797561761569935E-2,
                            4.76367800457137231464E-2, 2.07448227648435975150E-1, 4.94214826801497100753E-1,
                            9.99999999999999996796E-1};
static const double GQ[] = {-2.31581873324120129819E-5, 5.39605580493303397842E-4, -4.45641913851797240494E-3,
                            1.18139785222060435552E-2,  3.58236398605498653373E-2, -2.34591795718243348568E-1,
                            7.14304917030273074085E-2,  1.00000000000000000320E0};
static const double STIR[] = {7.87311395793093628397E-4, -2.29549961613378126380E-4, -2.68132617805781232825E-3,
                              3.47222221605458667310E-3, 8.33333333333482257126E-2};
static const double LA[] = {8.11614167470508450300E-4, -5.95061904284301438324E-4, 7.93650340457716943945E-4,
                            -2.77777777730099687205E-3, 8.33333333333331927722E-2};
static const double LB[] = {-1.37825152569120859100E3, -3.88016315134637840924E4, -3.31612992738871184744E5,
                            -1.16237097492762307383E6, -1.72173700820839662146E6, -8.53555664245765465627E5};
static const double LC[] = {-3.51815701436523470549E2, -1.70642106651881159223E4, -2.20528590553854454839E5,
                            -1.13933444367982507207E6, -2.53252307177582951285E6, -2.01889141433532773231E6};

static double stirf(double x) {
    double y, w, v;
    if (x >= MAXGAM) return INFINITY;
    w = 1.0 / x;
    w = 1.0 + w * polevl(w, STIR, 4);
    y = exp(x);
    if (x > MAXSTIR) {
        v = pow(x, 0.5 * x - 0.25);
        y = v * (v / y);
    } else {
        y = pow(x, x - 0.5) / y;
    }
    return SQTPI * y * w;
}

/* positive arguments only are reachable from bdtrc (a = k+1 >= 2, b = n-k >= 1) */
static double cephes_Gamma(double x) {
    double p, q, z;
    if (!isfinite(x)) return x;
    q = fabs(x);
    if (q > 33.0) {
        if (x < 0.0) return NAN; /* unreachable from bdtrc */
        return stirf(x);
    }
    z = 1.0;
    while (x >= 3.0) {
        x -= 1.0;
        z *= x;
    }
    while (x < 0.0) {
        if (x > -1.E-9) goto small;
        z /= x;
        x += 1.0;
    }
    while (x < 2.0) {
        if (x < 1.e-9) goto small;
        z /= x;
        x += 1.0;
    }
    if (x == 2.0) return z;
    x -= 2.0;
    p = polevl(x, GP, 6);
    q = polevl(x, GQ, 7);
    return z * p / q;
small:
    if (x == 0.0) return INFINITY;
    return z / ((1.0 + 0.5772156649015329 * x) * x);
}

double oracle_lgam(double x) {
    double p, q, u, w, z;
    if (!isfinite(x)) return x;
    if (x < -34.0) return NAN; /* unreachable from bdtrc */
    if (x < 13.0) {
        z = 1.0;
        p = 0.0;
        u = x;
        while (u >= 3.0) {
            p -= 1.0;
            u = x + p;
            z *= u;
        }
        while (u < 2.0) {
            if (u == 0.0) return INFINITY;
            z /= u;
            p += 1.0;
            u = x + p;
        }
        if (z < 0.0) z = -z;
        if (u == 2.0) return log(z);
        p -= 2.0;
        x = x + p;
        p = x * polevl(x, LB, 5) / p1evl(x, LC, 6);
        return log(z) + p;
    }
    if (x > MAXLGM) return INFINITY;
    q = (x - 0.5) * log(x) - x + LS2PI;
    if (x > 1.0e8) return q;
    p = 1.0 / (x * x);
    if (x >= 1000.0)
        q += ((7.9365079365079365079365e-4 * p - 2.7777777777777777777778e-3) * p + 0.0833333333333333333333) / x;
    else
        q += polevl(p, LA, 4) / x;
    (void)w;
    return q;
}

static double lbeta_asymp(double a, double b) {
    double r = oracle_lgam(b);
    r -= b * log(a);
    r += b * (1 - b) / (2 * a);
    r += b * (1 - b) * (1 - 2 * b) / (12 * a * a);
    r += -b * b * (1 - b) * (1 - b) / (12 * a * a * a);
    return r;
}

double oracle_lbeta(double a, double b) {
    double y;
    if (fabs(a) < fabs(b)) {
        y = a; a = b; b = y;
    }
    if (fabs(a) > ASYMP_FACTOR * fabs(b) && a > ASYMP_FACTOR) return lbeta_asymp(a, b);
    y = a + b;
    if (fabs(y) > MAXGAM || fabs(a) > MAXGAM || fabs(b) > MAXGAM) {
        y = oracle_lgam(y);
        y = oracle_lgam(b) - y;
        y = oracle_lgam(a) + y;
        return y;
    }
    y = cephes_Gamma(y);
    a = cephes_Gamma(a);
    b = cephes_Gamma(b);
    if (y == 0.0) return INFINITY;
    if (fabs(fabs(a) - fabs(y)) > fabs(fabs(b) - fabs(y))) {
        y = b / y;
        y *= a;
    } else {
        y = a / y;
        y *= b;
    }
    if (y < 0) y = -y;
    return log(y);
}

static double cephes_beta(double a, double b) {
    double y;
    if (fabs(a) < fabs(b)) {
        y = a; a = b; b = y;
    }
    if (fabs(a) > ASYMP_FACTOR * fabs(b) && a > ASYMP_FACTOR) return exp(lbeta_asymp(a, b));
    y = a + b;
    if (fabs(y) > MAXGAM || fabs(a) > MAXGAM || fabs(b) > MAXGAM) {
        y = oracle_lgam(y);
        y = oracle_lgam(b) - y;
        y = oracle_lgam(a) + y;
        if (y > MAXLOG) return INFINITY;
        return exp(y);
    }
    y = cephes_Gamma(y);
    a = cephes_Gamma(a);
    b = cephes_Gamma(b);
    if (y == 0.0) return INFINITY;
    if (fabs(fabs(a) - fabs(y)) > fabs(fabs(b) - fabs(y))) {
        y = b / y;
        y *= a;
    } else {
        y = a / y;
        y *= b;
    }
    return y;
}

/* ---- incomplete beta (cephes incbet.c) ---- */
static __thread int g_last_iters; /* diagnostics for tests: CF iterations of the last call (not thread-safe; tests only) */

static double incbcf(double a, double b, double x) {
    double xk, pk, pkm1, pkm2, qk, qkm1, qkm2;
    double k1 = a, k2 = a + b, k3 = a, k4 = a + 1.0, k5 = 1.0, k6 = b - 1.0, k7 = a + 1.0, k8 = a + 2.0;
    double r = 1.0, t, ans = 1.0, thresh = 3.0 * MACHEP;
    int n = 0;
    pkm2 = 0.0; qkm2 = 1.0; pkm1 = 1.0; qkm1 = 1.0;
    do {
        xk = -(x * k1 * k2) / (k3 * k4);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;

        xk = (x * k5 * k6) / (k7 * k8);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;

        if (qk != 0) r = pk / qk;
        if (r != 0) {
            t = fabs((ans - r) / r);
            ans = r;
        } else {
            t = 1.0;
        }
        if (t < thresh) break;

        k1 += 1.0; k2 += 1.0; k3 += 2.0; k4 += 2.0; k5 += 1.0; k6 -= 1.0; k7 += 2.0; k8 += 2.0;

        if ((fabs(qk) + fabs(pk)) > BIG) {
            pkm2 *= BIGINV; pkm1 *= BIGINV; qkm2 *= BIGINV; qkm1 *= BIGINV;
        }
        if ((fabs(qk) < BIGINV) || (fabs(pk) < BIGINV)) {
            pkm2 *= BIG; pkm1 *= BIG; qkm2 *= BIG; qkm1 *= BIG;
        }
    } while (++n < 300);
    g_last_iters = n;
    return ans;
}

static double incbd(double a, double b, double x) {
    double xk, pk, pkm1, pkm2, qk, qkm1, qkm2;
    double k1 = a, k2 = b - 1.0, k3 = a, k4 = a + 1.0, k5 = 1.0, k6 = a + b, k7 = a + 1.0, k8 = a + 2.0;
    double r = 1.0, t, ans = 1.0, z = x / (1.0 - x), thresh = 3.0 * MACHEP;
    int n = 0;
    pkm2 = 0.0; qkm2 = 1.0; pkm1 = 1.0; qkm1 = 1.0;
    do {
        xk = -(z * k1 * k2) / (k3 * k4);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;

        xk = (z * k5 * k6) / (k7 * k8);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;

        if (qk != 0) r = pk / qk;
        if (r != 0) {
            t = fabs((ans - r) / r);
            ans = r;
        } else {
            t = 1.0;
        }
        if (t < thresh) break;

        k1 += 1.0; k2 -= 1.0; k3 += 2.0; k4 += 2.0; k5 += 1.0; k6 += 1.0; k7 += 2.0; k8 += 2.0;

        if ((fabs(qk) + fabs(pk)) > BIG) {
            pkm2 *= BIGINV; pkm1 *= BIGINV; qkm2 *= BIGINV; qkm1 *= BIGINV;
        }
        if ((fabs(qk) < BIGINV) || (fabs(pk) < BIGINV)) {
            pkm2 *= BIG; pkm1 *= BIG; qkm2 *= BIG; qkm1 *= BIG;
        }
    } while (++n < 300);
    g_last_iters = n;
    return ans;
}

static double pseries(double a, double b, double x) {
    double s, t, u, v, n, t1, z, ai;
    ai = 1.0 / a;
    u = (1.0 - b) * x;
    v = u / (a + 1.0);
    t1 = v;
    t = u;
    n = 2.0;
    s = 0.0;
    z = MACHEP * ai;
    while (fabs(v) > z) {
        u = (n - b) * x / n;
        t *= u;
        v = t / (a + n);
        s += v;
        n += 1.0;
    }
    s += t1;
    s += ai;
    u = a * log(x);
    if ((a + b) < MAXGAM && fabs(u) < MAXLOG) {
        t = 1.0 / cephes_beta(a, b);
        s = s * t * pow(x, a);
    } else {
        t = -oracle_lbeta(a, b) + u + log(s);
        if (t < MINLOG) s = 0.0;
        else s = exp(t);
    }
    return s;
}

double oracle_incbet(double aa, double bb, double xx) {
    double a, b, t, x, xc, w, y;
    int flag = 0;
    if (aa <= 0.0 || bb <= 0.0) return NAN;
    if ((xx <= 0.0) || (xx >= 1.0)) {
        if (xx == 0.0) return 0.0;
        if (xx == 1.0) return 1.0;
        return NAN;
    }
    if ((bb * xx) <= 1.0 && xx <= 0.95) return pseries(aa, bb, xx);
    w = 1.0 - xx;
    if (xx > (aa / (aa + bb))) {
        flag = 1; a = bb; b = aa; xc = xx; x = w;
    } else {
        a = aa; b = bb; xc = w; x = xx;
    }
    if (flag == 1 && (b * x) <= 1.0 && x <= 0.95) {
        t = pseries(a, b, x);
        goto done;
    }
    y = x * (a + b - 2.0) - (a - 1.0);
    if (y < 0.0) w = incbcf(a, b, x);
    else w = incbd(a, b, x) / xc;

    y = a * log(x);
    t = b * log(xc);
    if ((a + b) < MAXGAM && fabs(y) < MAXLOG && fabs(t) < MAXLOG) {
        t = pow(xc, b);
        t *= pow(x, a);
        t /= a;
        t *= w;
        t *= 1.0 / cephes_beta(a, b);
        goto done;
    }
    y += t - oracle_lbeta(a, b);
    y += log(w / a);
    if (y < MINLOG) t = 0.0;
    else t = exp(y);
done:
    if (flag == 1) {
        if (t <= MACHEP) t = 1.0 - MACHEP;
        else t = 1.0 - t;
    }
    return t;
}

/* cephes bdtrc(k, n, p) with scipy's 'dld->d' wrapper: n arrives as a C long and is cast to int (int32 wrap,
 * SURVEY F5); callers keep n < 2^31. */
double oracle_bdtrc(double k, int64_t n_in, double p) {
    int n = (int)n_in;
    double dk, dn, fk;
    if (isnan(p) || isnan(k)) return NAN;
    if (p < 0.0 || p > 1.0) return NAN;
    fk = floor(k);
    if (fk < 0) return 1.0;
    if (n < fk) return NAN;
    if (fk == n) return 0.0;
    dn = n - fk;
    if (k == 0) {
        if (p < .01) dk = -expm1(dn * log1p(-p));
        else dk = 1.0 - pow(1.0 - p, dn);
    } else {
        dk = fk + 1;
        dk = oracle_incbet(dk, dn, p);
    }
    return dk;
}

/* vector entry: out[i] = bdtrc(k[i], n, p[i]); OpenMP over i (the cpu_baseline leg states the thread count) */
void oracle_bdtrc_vec(const double *k, int64_t n, const double *p, double *out, int64_t len) {
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < len; i++) out[i] = oracle_bdtrc(k[i], n, p[i]);
}

/* same, and also report the continued-fraction iteration count per element (0 when no CF ran); serial */
void oracle_bdtrc_vec_iters(const double *k, int64_t n, const double *p, double *out, int32_t *iters, int64_t len) {
    for (int64_t i = 0; i < len; i++) {
        g_last_iters = 0;
        out[i] = oracle_bdtrc(k[i], n, p[i]);
        iters[i] = g_last_iters;
    }
}
