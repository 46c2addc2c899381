// Smoothing-spline fit of fit_Spline's first stage on the host: `UnivariateSpline(x, y, s=min(y)^2)` (reference
// fithic/fithic.py:951) is scipy's FITPACK `curfit` (P. Dierckx: fpcurf / fpknot / fpgivs / fprota / fpback / fpdisc /
// fprati / fpbspl; third party, not under /root/reference), called through scipy.interpolate._fitpack2._curfit and, when
// the first call ends with ier = 1 ("nest too small"), once more with nest = m + k + 1 (UnivariateSpline._reset_nest).
//
// The fit sits on the critical path of every spline pass and of every rank (the GPU idles while <= noOfBins points are
// fitted), and the two scipy calls cost 0.6 ms of a 2 ms step at 8 GPUs.  This is a restatement of Dierckx's published
// algorithm with the SAME sequence of IEEE operations (no contraction: the file is compiled with -ffp-contract=off), so
// knots and coefficients equal scipy's bit for bit (tests/test_host.py pins it against scipy on the golden fixtures and
// on random inputs).  What is NOT the same is the control flow around it: scipy's second call starts again from the
// polynomial and repeats every least-squares fit of the first one; here the state at the moment the first call runs into
// its storage limit is kept and the second call resumes from there (identical arithmetic from that point on).
//
// Arrays are 1-based inside (index 0 unused) so that the recurrences read like the published ones.
#pragma once
#include <math.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <vector>

namespace fhc {
namespace fitpack {

struct Work {
    int m = 0, nest = 0, k = 0;
    std::vector<double> fpint, z, a, b, g, q, t, c;
    std::vector<int> nrdata, row_l, col_time, op_time, op_list, bucket;
    std::vector<double> row_h, row_y, row_knots;
    bool row_knots_valid = false;
    void size(int m_, int nest_, int k_) {
        m = m_;
        nest = nest_;
        k = k_;
        const int k1 = k + 1, k2 = k + 2;
        fpint.assign((size_t)nest + 1, 0.0);
        z.assign((size_t)nest + 1, 0.0);
        a.assign((size_t)(nest + 1) * (k1 + 1), 0.0);
        b.assign((size_t)(nest + 1) * (k2 + 1), 0.0);
        g.assign((size_t)(nest + 1) * (k2 + 1), 0.0);
        q.assign((size_t)(m + 1) * (k1 + 1), 0.0);
        t.assign((size_t)nest + 1, 0.0);
        c.assign((size_t)nest + 1, 0.0);
        nrdata.assign((size_t)nest + 1, 0);
        const size_t rows = (size_t)(m > nest ? m : nest) + 2;
        row_l.assign(rows, 0);
        row_h.assign(rows * 8, 0.0);
        row_y.assign(rows, 0.0);
        row_knots.assign(rows * 8, 0.0);
        row_knots_valid = false;
        col_time.assign((size_t)nest + 2, 0);
        op_time.assign((size_t)m * (k1 + 1) + 8, 0);
        op_list.assign((size_t)m * (k1 + 1) + 8, 0);
        bucket.assign((size_t)m * (k1 + 1) + 16, 0);
    }
};

// fpgivs: parameters of a givens rotation
static inline void fpgivs(double piv, double &ww, double &cs, double &sn) {
    const double store = fabs(piv);
    double dd;
    if (store >= ww) {
        const double r = ww / piv;
        dd = store * sqrt(1.0 + r * r);
    } else {
        const double r = piv / ww;
        dd = ww * sqrt(1.0 + r * r);
    }
    cs = ww / dd;
    sn = piv / dd;
    ww = dd;
}

// fprota: apply a givens rotation to a and b
static inline void fprota(double cs, double sn, double &a, double &b) {
    const double stor1 = a, stor2 = b;
    b = cs * stor2 + sn * stor1;
    a = cs * stor1 - sn * stor2;
}

// fpbspl: the k + 1 non-zero b-splines of degree k at t(l) <= x < t(l+1) (de Boor / Cox)
static inline void fpbspl(const double *t, int k, double x, int l, double *h /*1-based, k+1*/) {
    double hh[8];
    h[1] = 1.0;
    for (int j = 1; j <= k; ++j) {
        for (int i = 1; i <= j; ++i) hh[i] = h[i];
        h[1] = 0.0;
        for (int i = 1; i <= j; ++i) {
            const int li = l + i, lj = li - j;
            if (t[li] == t[lj]) {
                h[i + 1] = 0.0;
                continue;
            }
            const double f = hh[i] / (t[li] - t[lj]);
            h[i] = h[i] + f * (t[li] - x);
            h[i + 1] = f * (x - t[lj]);
        }
    }
}

// fpback: backward substitution for an n x n upper triangular band matrix of bandwidth k, stored in a(nest+1, k+1)
static inline void fpback(const double *a, int lda, const double *z, int n, int k, double *c) {
#define FHC_A(i, j) a[(size_t)(j) * lda + (i)]
    const int k1 = k - 1;
    c[n] = z[n] / FHC_A(n, 1);
    int i = n - 1;
    if (i == 0) return;
    for (int j = 2; j <= n; ++j) {
        double store = z[i];
        int i1 = k1;
        if (j <= k1) i1 = j - 1;
        int mm = i;
        for (int l = 1; l <= i1; ++l) {
            mm += 1;
            store = store - c[mm] * FHC_A(i, l + 1);
        }
        c[i] = store / FHC_A(i, 1);
        i -= 1;
    }
#undef FHC_A
}

// fpdisc: discontinuity jumps of the k-th derivative of the b-splines at the interior knots
static inline void fpdisc(const double *t, int n, int k2, double *b, int ldb) {
#define FHC_B(i, j) b[(size_t)(j) * ldb + (i)]
    double h[16];
    const int k1 = k2 - 1, k = k1 - 1, nk1 = n - k1, nrint = nk1 - k;
    const double an = (double)nrint;
    const double fac = an / (t[nk1 + 1] - t[k1]);
    for (int l = k2; l <= nk1; ++l) {
        const int lmk = l - k1;
        for (int j = 1; j <= k1; ++j) {
            const int ik = j + k1, lj = l + j, lk = lj - k2;
            h[j] = t[l] - t[lk];
            h[ik] = t[l] - t[lj];
        }
        int lp = lmk;
        for (int j = 1; j <= k2; ++j) {
            int jk = j;
            double prod = h[j];
            for (int i = 1; i <= k; ++i) {
                jk += 1;
                prod = prod * h[jk] * fac;
            }
            const int lk = lp + k1;
            FHC_B(lmk, j) = (t[lk] - t[lp]) / prod;
            lp += 1;
        }
    }
#undef FHC_B
}

// fprati: root of the rational interpolant through (p1, f1), (p2, f2), (p3, f3); p3 < 0 stands for infinity
static inline double fprati(double &p1, double &f1, double p2, double f2, double &p3, double &f3) {
    double p;
    if (p3 > 0.0) {
        const double h1 = f1 * (f2 - f3), h2 = f2 * (f3 - f1), h3 = f3 * (f1 - f2);
        p = -(p1 * p2 * h3 + p2 * p3 * h1 + p3 * p1 * h2) / (p1 * h1 + p2 * h2 + p3 * h3);
    } else {
        p = (p1 * (f1 - f3) * f2 - p2 * (f2 - f3) * f1) / ((f1 - f2) * f3);
    }
    if (f2 < 0.0) {
        p3 = p2;
        f3 = f2;
    } else {
        p1 = p2;
        f1 = f2;
    }
    return p;
}

// fpknot: one more knot in the interval with the largest residual sum, at a data point in its middle
static inline void fpknot(const double *x, double *t, int &n, double *fpint, int *nrdata, int &nrint, int istart) {
    const int k = (n - nrint - 1) / 2;
    double fpmax = 0.0;
    int jbegin = istart, number = 0, maxpt = 0, maxbeg = 0;
    for (int j = 1; j <= nrint; ++j) {
        const int jpoint = nrdata[j];
        if (!(fpmax >= fpint[j] || jpoint == 0)) {
            fpmax = fpint[j];
            number = j;
            maxpt = jpoint;
            maxbeg = jbegin;
        }
        jbegin = jbegin + jpoint + 1;
    }
    if (number == 0) {
        // every interval with data has a zero residual sum: the published routine leaves `number` undefined here; take the
        // interval with the most points (what later FITPACK revisions do) so that the knot count still grows
        jbegin = istart;
        for (int j = 1; j <= nrint; ++j) {
            const int jpoint = nrdata[j];
            if (jpoint > maxpt) {
                number = j;
                maxpt = jpoint;
                maxbeg = jbegin;
            }
            jbegin = jbegin + jpoint + 1;
        }
        if (number == 0) {  // no interval holds a data point: nothing can be added
            n += 1;         // keep the caller's loop finite; t gets a repeated knot at the right end
            nrint += 1;
            return;
        }
    }
    const int ihalf = maxpt / 2 + 1;
    const int nrx = maxbeg + ihalf;
    const int next = number + 1;
    if (next <= nrint) {
        for (int j = next; j <= nrint; ++j) {
            const int jj = next + nrint - j;
            fpint[jj + 1] = fpint[jj];
            nrdata[jj + 1] = nrdata[jj];
            const int jk = jj + k;
            t[jk + 1] = t[jk];
        }
    }
    nrdata[number] = ihalf - 1;
    nrdata[next] = maxpt - ihalf;
    const double am = (double)maxpt;
    double an = (double)nrdata[number];
    fpint[number] = fpmax * an / am;
    an = (double)nrdata[next];
    fpint[next] = fpmax * an / am;
    const int jk = next + k;
    t[jk] = x[nrx];
    n += 1;
    nrint += 1;
}

// State of the knot search at the moment a run with storage limit `nest_small` stops adding knots (n == nest_small inside
// a batch of nplus knots): everything a run with a larger limit needs to go on from there.
struct Resume {
    bool valid = false;
    int n = 0, nrint = 0, nplus = 0, l_next = 0, iter = 0;
    double fpold = 0.0, fp0 = 0.0;
    std::vector<double> t, fpint;
    std::vector<int> nrdata;
};

// fpcurf with iopt = 0, weights 1, xb = x(1), xe = x(m).  x, y: 1-based arrays of m points, x ascending.
// On return W.t(1..n), W.c(1..n) hold knots and coefficients.  `save`: filled when the run stops adding knots because
// n == nest (for a later run with a larger nest); `from`: resume such a state instead of starting from the polynomial.
static inline int fpcurf(const double *x, const double *y, int m, int k, double s, int nest, Work &W, int &n_out,
                         double &fp_out, Resume *save, const Resume *from) {
    const double tol = 0.001;
    const int maxit = 20;
    const int k1 = k + 1, k2 = k + 2;
    const double xb = x[1], xe = x[m];
    const double con1 = 0.1, con9 = 0.9, con4 = 0.04, half = 0.5;
    double *t = W.t.data(), *c = W.c.data(), *fpint = W.fpint.data(), *z = W.z.data();
    int *nrdata = W.nrdata.data();
    const int lda = nest + 1;
#define FHC_A(i, j) W.a[(size_t)(j) * lda + (i)]
#define FHC_G(i, j) W.g[(size_t)(j) * lda + (i)]
#define FHC_BB(i, j) W.b[(size_t)(j) * lda + (i)]
#define FHC_Q(i, j) W.q[(size_t)(j) * (m + 1) + (i)]
    double h[8];
    const int nmin = 2 * k1;
    const double acc = tol * s;
    const int nmax = m + k1;
    int n = 0, ier = 0, nplus = 0, nrint = 0, nk1 = 0;
    double fp = 0.0, fp0 = 0.0, fpold = 0.0, fpms = 0.0;
    bool interp_knots = false;
    int resume_l = 0;  // > 0: enter the knot-adding batch at this position (state restored from `from`)
    int iter0 = 1;

    if (from != nullptr && from->valid) {
        n = from->n;
        nrint = from->nrint;
        nplus = from->nplus;
        fpold = from->fpold;
        fp0 = from->fp0;
        for (int i = 1; i <= n; ++i) t[i] = from->t[(size_t)i];
        for (size_t i = 1; i < from->fpint.size() && i <= (size_t)nest; ++i) fpint[i] = from->fpint[i];
        for (size_t i = 1; i < from->nrdata.size() && i <= (size_t)nest; ++i) nrdata[i] = from->nrdata[i];
        resume_l = from->l_next;
        iter0 = from->iter;
    } else if (s > 0.0) {
        n = nmin;
        fpold = 0.0;
        nplus = 0;
        nrdata[1] = m - 2;
    } else {
        // s = 0: interpolating spline
        n = nmax;
        if (nmax > nest) {
            n_out = n;
            fp_out = fp;
            return 1;
        }
        interp_knots = true;
    }

    bool accepted = false;  // reached label 250 (knots fixed, smoothing spline follows)
    for (;;) {              // label 10 / 60: (re)start of the main loop
        if (interp_knots) {
            const int mk1 = m - k1;
            if (mk1 != 0) {
                const int k3 = k / 2;
                int i = k2, j = k3 + 2;
                if (k3 * 2 == k) {
                    for (int l = 1; l <= mk1; ++l) {
                        t[i] = (x[j] + x[j - 1]) * half;
                        i += 1;
                        j += 1;
                    }
                } else {
                    for (int l = 1; l <= mk1; ++l) {
                        t[i] = x[j];
                        i += 1;
                        j += 1;
                    }
                }
            }
            interp_knots = false;
        }
        bool restart = false;
        for (int iter = iter0; iter <= m; ++iter) {
            if (resume_l == 0) {
#ifdef FHC_FIT_STATS
                g_lsq++;
                const double t_lsq0 = wall_ms();
#endif
                if (n == nmin) ier = -2;
                nrint = n - nmin + 1;
                nk1 = n - k1;
                {
                    int i = n;
                    for (int j = 1; j <= k1; ++j) {
                        t[j] = xb;
                        t[i] = xe;
                        i -= 1;
                    }
                }
                fp = 0.0;
                for (int i = 1; i <= nk1; ++i) {
                    z[i] = 0.0;
                    for (int j = 1; j <= k1; ++j) FHC_A(i, j) = 0.0;
                }
                // The rows of the observation matrix are rotated into the triangle one after the other in the published
                // routine; a rotation only has to wait for the earlier rotations on the SAME column and for the previous
                // step of its own row, so the steps are executed in wavefronts of mutually independent rotations (same
                // operations on the same operands, hence the same bits; the divisions and square roots of a wavefront
                // overlap in the pipeline instead of queueing behind each other).
                {
                    int l = k1;
                    for (int it = 1; it <= m; ++it) {
                        const double xi = x[it];
                        while (!(xi < t[l + 1] || l == nk1)) l += 1;
                        W.row_l[(size_t)it] = l;
                        // the b-splines at x(it) depend on the 2k knots around its interval only: when those are the ones
                        // of the previous knot set (one more knot changes a handful of rows), the stored values are reused
                        double *kn = &W.row_knots[(size_t)it * 8];
                        bool same = W.row_knots_valid;
                        for (int u = 0; u < 2 * k && same; ++u) same = kn[u] == t[l - k + 1 + u];
                        if (same) {
                            for (int i = 1; i <= k1; ++i) h[i] = FHC_Q(it, i);
                        } else {
                            fpbspl(t, k, xi, l, h);
                            for (int u = 0; u < 2 * k; ++u) kn[u] = t[l - k + 1 + u];
                        }
                        double *hr = &W.row_h[(size_t)it * 8];
                        for (int i = 1; i <= k1; ++i) {
                            FHC_Q(it, i) = h[i];
                            hr[i] = h[i] * 1.0;
                        }
                        W.row_y[(size_t)it] = y[it] * 1.0;
                    }
#ifdef FHC_FIT_STATS
                    g_ta += wall_ms() - t_lsq0;
                    const double t_s0 = wall_ms();
#endif
                    W.row_knots_valid = true;
                    // schedule: time of step (it, i) = 1 + max(time of (it, i - 1), last time column j was touched)
                    const int nops = m * k1;
                    for (int j = 0; j <= nk1; ++j) W.col_time[(size_t)j] = 0;
                    int tmax = 0;
                    for (int it = 1; it <= m; ++it) {
                        int tr = 0;
                        int j = W.row_l[(size_t)it] - k1;
                        for (int i = 1; i <= k1; ++i) {
                            j += 1;
                            const int ct = W.col_time[(size_t)j];
                            tr = (tr > ct ? tr : ct) + 1;
                            W.col_time[(size_t)j] = tr;
                            W.op_time[(size_t)((it - 1) * k1 + (i - 1))] = tr;
                            if (tr > tmax) tmax = tr;
                        }
                    }
                    for (int tt = 0; tt <= tmax + 1; ++tt) W.bucket[(size_t)tt] = 0;
                    for (int o = 0; o < nops; ++o) W.bucket[(size_t)W.op_time[(size_t)o] + 1] += 1;
                    for (int tt = 1; tt <= tmax + 1; ++tt) W.bucket[(size_t)tt] += W.bucket[(size_t)tt - 1];
                    for (int o = 0; o < nops; ++o) W.op_list[(size_t)W.bucket[(size_t)W.op_time[(size_t)o]]++] = o;
                    // (bucket[tt] now marks the end of wavefront tt; the ops of a row/column stay in row order inside it)
#ifdef FHC_FIT_STATS
                    g_tb += wall_ms() - t_s0;
#endif
                    int pos = 0;
                    auto op1 = [&](int o) {
                        const int it = o / k1 + 1, i = o % k1 + 1;
                        double *hr = &W.row_h[(size_t)it * 8];
                        const double piv = hr[i];
                        if (piv == 0.0) return;
                        const int j = W.row_l[(size_t)it] - k1 + i;
                        double cs, sn;
                        fpgivs(piv, FHC_A(j, 1), cs, sn);
                        fprota(cs, sn, W.row_y[(size_t)it], z[j]);
                        int i2 = 1;
                        for (int i1 = i + 1; i1 <= k1; ++i1) {
                            i2 += 1;
                            fprota(cs, sn, hr[i1], FHC_A(j, i2));
                        }
                    };
                    for (int tt = 1; tt <= tmax; ++tt) {
                        const int end = W.bucket[(size_t)tt];
                        for (; pos < end; ++pos) op1(W.op_list[(size_t)pos]);
                    }
                    for (int it = 1; it <= m; ++it) fp = fp + W.row_y[(size_t)it] * W.row_y[(size_t)it];
                }
                if (ier == -2) fp0 = fp;
                fpint[n] = fp0;
                fpint[n - 1] = fpold;
                nrdata[n] = nplus;
                fpback(W.a.data(), lda, z, nk1, k1, c);
#ifdef FHC_FIT_STATS
                g_t1 += wall_ms() - t_lsq0;
#endif
                fpms = fp - s;
                if (fabs(fpms) < acc) goto done;
                if (fpms < 0.0) {
                    accepted = true;
                    break;
                }
                if (n == nmax) {
                    ier = -1;
                    goto done;
                }
                if (n == nest) {
                    ier = 1;
                    goto done;
                }
                if (ier == 0) {
                    int npl1 = nplus * 2;
                    const double rn = (double)nplus;
                    if (fpold - fp > acc) npl1 = (int)(rn * fpms / (fpold - fp));
                    int mx = npl1;
                    if (nplus / 2 > mx) mx = nplus / 2;
                    if (1 > mx) mx = 1;
                    nplus = nplus * 2 < mx ? nplus * 2 : mx;
                } else {
                    nplus = 1;
                    ier = 0;
                }
                fpold = fp;
                double fpart = 0.0;
                int i = 1;
                int l = k2;
                int nw = 0;
                for (int it = 1; it <= m; ++it) {
                    if (!(x[it] < t[l] || l > nk1)) {
                        nw = 1;
                        l += 1;
                    }
                    double term = 0.0;
                    int l0 = l - k2;
                    for (int j = 1; j <= k1; ++j) {
                        l0 += 1;
                        term = term + c[l0] * FHC_Q(it, j);
                    }
                    const double r = 1.0 * (term - y[it]);
                    term = r * r;
                    fpart = fpart + term;
                    if (nw == 0) continue;
                    const double store = term * half;
                    fpint[i] = fpart - store;
                    i += 1;
                    fpart = store;
                    nw = 0;
                }
                fpint[nrint] = fpart;
            }
            {
                const int lfirst = resume_l > 0 ? resume_l : 1;
                resume_l = 0;
                bool to_interp = false;
                for (int l = lfirst; l <= nplus; ++l) {
                    fpknot(x, t, n, fpint, nrdata, nrint, 1);
                    if (n == nmax) {
                        to_interp = true;
                        break;
                    }
                    if (n == nest) {
                        if (save != nullptr) {
                            // a run with more storage would go on from here (with the rest of this batch, if any)
                            save->valid = true;
                            save->n = n;
                            save->nrint = nrint;
                            save->nplus = nplus;
                            save->l_next = l + 1;
                            save->iter = iter;
                            save->fpold = fpold;
                            save->fp0 = fp0;
                            save->t.assign(t, t + nest + 1);
                            save->fpint.assign(fpint, fpint + nest + 1);
                            save->nrdata.assign(nrdata, nrdata + nest + 1);
                        }
                        break;
                    }
                }
                if (to_interp) {
                    interp_knots = true;
                    restart = true;
                    break;
                }
            }
        }
        if (restart) {
            iter0 = 1;
            continue;
        }
        break;
    }
    (void)accepted;
    // label 250: test whether the least-squares polynomial is a solution
    if (ier == -2) goto done;
    {
        // part 2: the smoothing spline sp(x) for the knots found
        fpdisc(t, n, k2, W.b.data(), lda);
        double p1 = 0.0, f1 = fp0 - s, p3 = -1.0, f3 = fpms, p = 0.0;
        for (int i = 1; i <= nk1; ++i) p = p + FHC_A(i, 1);
        const double rn = (double)nk1;
        p = rn / p;
        int ich1 = 0, ich3 = 0;
        const int n8 = n - nmin;
        for (int iter = 1; iter <= maxit; ++iter) {
#ifdef FHC_FIT_STATS
            g_pit++;
            const double t_p0 = wall_ms();
#endif
            const double pinv = 1.0 / p;
            for (int i = 1; i <= nk1; ++i) {
                c[i] = z[i];
                FHC_G(i, k2) = 0.0;
                for (int j = 1; j <= k1; ++j) FHC_G(i, j) = FHC_A(i, j);
            }
            // the rows of b (weight 1 / p) are rotated into the triangle: row `it` works on column j at time it + j, after
            // row it - 1 has left that column -- the rotations of one time step are independent (see the note above), and
            // two of them at a time go through the SSE2 unit (same IEEE operations per lane; division and square root are
            // what the step costs, and the packed forms have twice the throughput).  Row state is kept transposed and in
            // reverse row order so that the rows it, it - 1 and their columns j, j + 1 are both adjacent in memory.
            {
                const int hs = n8 + 2;  // stride of the transposed row state: hT[i * hs + (n8 - it)]
                double *hT = W.row_h.data();
                double *yT = W.row_y.data();
                for (int it = 1; it <= n8; ++it) {
                    const int r = n8 - it;
                    for (int i = 1; i <= k2; ++i) hT[(size_t)i * hs + r] = FHC_BB(it, i) * pinv;
                    yT[r] = 0.0;
                }
                auto step1 = [&](int it, int j) {
                    const int r = n8 - it;
                    const double piv = hT[(size_t)1 * hs + r];
                    double cs, sn;
                    fpgivs(piv, FHC_G(j, 1), cs, sn);
                    fprota(cs, sn, yT[r], c[j]);
                    if (j == nk1) return;
                    int i2 = k1;
                    if (j > n8) i2 = nk1 - j;
                    for (int i = 1; i <= i2; ++i) {
                        const int i1 = i + 1;
                        double hv = hT[(size_t)i1 * hs + r];
                        fprota(cs, sn, hv, FHC_G(j, i1));
                        hT[(size_t)i1 * hs + r] = hv;
                        hT[(size_t)i * hs + r] = hv;
                    }
                    hT[(size_t)(i2 + 1) * hs + r] = 0.0;
                };
                for (int tau = 2; tau <= n8 + nk1; ++tau) {
                    const int lo = tau - nk1 > 1 ? tau - nk1 : 1;
                    const int hi = tau / 2 < n8 ? tau / 2 : n8;
                    int it = hi;
#if defined(__SSE2__)
                    for (; it - 1 >= lo; it -= 2) {
                        const int j = tau - it;  // lane 0: row it, column j; lane 1: row it - 1, column j + 1
                        if (j + 1 >= nk1 || j + 1 > n8) break;  // the last columns have shorter rows: one at a time
                        const int r = n8 - it;
                        const __m128d sign = _mm_set1_pd(-0.0), one = _mm_set1_pd(1.0);
                        const __m128d piv = _mm_loadu_pd(&hT[(size_t)1 * hs + r]);
                        const __m128d ww = _mm_loadu_pd(&FHC_G(j, 1));
                        const __m128d store = _mm_andnot_pd(sign, piv);
                        const __m128d ge = _mm_cmpge_pd(store, ww);
                        const __m128d num = _mm_or_pd(_mm_and_pd(ge, ww), _mm_andnot_pd(ge, piv));
                        const __m128d den = _mm_or_pd(_mm_and_pd(ge, piv), _mm_andnot_pd(ge, ww));
                        const __m128d base = _mm_or_pd(_mm_and_pd(ge, store), _mm_andnot_pd(ge, ww));
                        const __m128d rr = _mm_div_pd(num, den);
                        const __m128d dd = _mm_mul_pd(base, _mm_sqrt_pd(_mm_add_pd(one, _mm_mul_pd(rr, rr))));
                        const __m128d cs = _mm_div_pd(ww, dd), sn = _mm_div_pd(piv, dd);
                        _mm_storeu_pd(&FHC_G(j, 1), dd);
                        {
                            const __m128d a0 = _mm_loadu_pd(&yT[r]), b0 = _mm_loadu_pd(&c[j]);
                            _mm_storeu_pd(&c[j], _mm_add_pd(_mm_mul_pd(cs, b0), _mm_mul_pd(sn, a0)));
                            _mm_storeu_pd(&yT[r], _mm_sub_pd(_mm_mul_pd(cs, a0), _mm_mul_pd(sn, b0)));
                        }
                        for (int i = 1; i <= k1; ++i) {
                            const int i1 = i + 1;
                            const __m128d a0 = _mm_loadu_pd(&hT[(size_t)i1 * hs + r]), b0 = _mm_loadu_pd(&FHC_G(j, i1));
                            _mm_storeu_pd(&FHC_G(j, i1), _mm_add_pd(_mm_mul_pd(cs, b0), _mm_mul_pd(sn, a0)));
                            _mm_storeu_pd(&hT[(size_t)i * hs + r], _mm_sub_pd(_mm_mul_pd(cs, a0), _mm_mul_pd(sn, b0)));
                        }
                        _mm_storeu_pd(&hT[(size_t)k2 * hs + r], _mm_setzero_pd());
                    }
#endif
                    for (; it >= lo; --it) step1(it, tau - it);
                }
            }
            fpback(W.g.data(), lda, c, nk1, k2, c);
#ifdef FHC_FIT_STATS
            g_t2 += wall_ms() - t_p0;
#endif
            fp = 0.0;
            int l = k2;
            for (int it = 1; it <= m; ++it) {
                if (!(x[it] < t[l] || l > nk1)) l += 1;
                int l0 = l - k2;
                double term = 0.0;
                for (int j = 1; j <= k1; ++j) {
                    l0 += 1;
                    term = term + c[l0] * FHC_Q(it, j);
                }
                const double r = 1.0 * (term - y[it]);
                fp = fp + r * r;
            }
            fpms = fp - s;
            if (fabs(fpms) < acc) goto done;
            if (iter == maxit) {
                ier = 3;
                goto done;
            }
            const double p2 = p, f2 = fpms;
            if (ich3 == 0) {
                if (!((f2 - f3) > acc)) {
                    p3 = p2;
                    f3 = f2;
                    p = p * con4;
                    if (p <= p1) p = p1 * con9 + p2 * con1;
                    continue;
                }
                if (f2 < 0.0) ich3 = 1;
            }
            if (ich1 == 0) {
                if (!((f1 - f2) > acc)) {
                    p1 = p2;
                    f1 = f2;
                    p = p / con4;
                    if (p3 < 0.0) continue;
                    if (p >= p3) p = p2 * con1 + p3 * con9;
                    continue;
                }
                if (f2 > 0.0) ich1 = 1;
            }
            if (f2 >= f1 || f2 <= f3) {
                ier = 2;
                goto done;
            }
            p = fprati(p1, f1, p2, f2, p3, f3);
        }
    }
done:
    n_out = n;
    fp_out = fp;
    return ier;
#undef FHC_A
#undef FHC_G
#undef FHC_BB
#undef FHC_Q
}

// UnivariateSpline(x, y, k = 3, s): curfit with nest = max(m / 2, 2 (k + 1)) (m + k + 1 when s == 0) and, if that ends
// with ier = 1, again with nest = m + k + 1.  x, y: 0-based arrays of m points.  t, c: room for m + k + 1 doubles each.
// Returns ier of the last run; *n_out = number of knots.
static inline int univariate_spline(const double *x0, const double *y0, int m, int k, double s, double *t, double *c,
                                    int *n_out, double *fp_out, int *calls_out) {
    std::vector<double> x((size_t)m + 1), y((size_t)m + 1);
    for (int i = 0; i < m; ++i) {
        x[(size_t)i + 1] = x0[i];
        y[(size_t)i + 1] = y0[i];
    }
    const int k1 = k + 1;
    const int nest_full = m + k + 1;
    int nest = s == 0.0 ? nest_full : (m / 2 > 2 * k1 ? m / 2 : 2 * k1);
    Work W;
    W.size(m, nest, k);
    Resume keep;
    int n = 0;
    double fp = 0.0;
    int ier = fpcurf(x.data(), y.data(), m, k, s, nest, W, n, fp, &keep, nullptr);
    int calls = 1;
    if (ier == 1 && nest < nest_full) {
        Work W2;
        W2.size(m, nest_full, k);
        if (keep.valid) {
            ier = fpcurf(x.data(), y.data(), m, k, s, nest_full, W2, n, fp, nullptr, &keep);
        } else {
            ier = fpcurf(x.data(), y.data(), m, k, s, nest_full, W2, n, fp, nullptr, nullptr);
        }
        calls = 2;
        for (int i = 0; i < n; ++i) {
            t[i] = W2.t[(size_t)i + 1];
            c[i] = W2.c[(size_t)i + 1];
        }
    } else {
        for (int i = 0; i < n; ++i) {
            t[i] = W.t[(size_t)i + 1];
            c[i] = W.c[(size_t)i + 1];
        }
    }
    *n_out = n;
    if (fp_out) *fp_out = fp;
    if (calls_out) *calls_out = calls;
    return ier;
}

}  // namespace fitpack
}  // namespace fhc
