// Host-side O(D) bookkeeping between K1 and K2: equal-occupancy binning and possible-pair counts.
//
// These two stages are sequential scans over at most D (~5e4) distances / sum(chromosome bins) (~6e5) slots with
// order-dependent integer and double accumulations; they run on the host between two kernels and must be bit exact,
// so they are plain C++ mirroring the reference's evaluation order:
//   fhc_host_make_bins   <- makeBinsFromInteractions   (reference fithic/fithic.py:463-553)
//   fhc_host_frag_pairs  <- generate_FragPairs, fixed-size branch (fithic/fithic.py:596-689)
#include <emmintrin.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include <math.h>

#include "cephes_dev.cuh"
#include "common.cuh"
#include "fitpack_host.cuh"

extern "C" int fhc_host_make_bins(const int64_t *dists, const int64_t *sums, int64_t m, int32_t noOfBins, int64_t N,
                                  int64_t *bin_lb, int64_t *bin_ub, int64_t *bin_sumcc) {
    FHC_REQUIRE(m >= 0 && noOfBins > 0, FHC_E_INVALID, "fhc_host_make_bins: need m >= 0 and noOfBins > 0");
    FHC_REQUIRE(m == 0 || (dists && sums), FHC_E_INVALID, "fhc_host_make_bins: null input");
    FHC_REQUIRE(bin_lb && bin_ub && bin_sumcc, FHC_E_INVALID, "fhc_host_make_bins: null output");
    // desiredPerBin = observedIntraInRangeSum / noOfBins  (true division, :476)
    double desired = (double)N / (double)noOfBins;
    int64_t total = 0;     // interactionTotalForBinTermination (:479)
    int64_t acc = 0;       // currentBinContactCount
    int64_t binsum = 0;    // sum of counts of the distances gathered in the open bin
    int64_t prev_ub = -1;  // last distance of the previous closed bin
    int nb = 0;            // closed bins

    for (int64_t i = 0; i < m; ++i) {
        const int64_t cc = sums[i];
        total += cc;
        bool full;
        // Python compares int with float exactly; both sides are < 2^53 here so the double compare is exact too
        if ((double)cc >= desired) {
            full = true;  // :481
        } else if ((double)(acc + cc) >= desired) {
            full = true;  // :486
        } else {
            full = false;
            acc += cc;  // :493
        }
        binsum += cc;

        if (full) {
            if (nb >= noOfBins) {
                // cannot happen for consistent input (sum of sums == N); refuse rather than write out of bounds
                fhc::set_error("fhc_host_make_bins: more than noOfBins=%d bins closed", noOfBins);
                return FHC_E_RANGE;
            }
            bin_lb[nb] = nb == 0 ? 0 : prev_ub + 1;  // :518-521
            bin_ub[nb] = dists[i];
            bin_sumcc[nb] = binsum;
            prev_ub = dists[i];
            nb += 1;
            if (nb < noOfBins) desired = 1.0 * (double)(N - total) / (double)(noOfBins - nb);  // :500-502
            acc = 0;
            binsum = 0;

        }
    }
    // distances after the last closed bin are dropped, as in the reference
    return nb;
}

extern "C" int fhc_host_frag_pairs(const int64_t *chr_n, const int64_t *chr_maxmid, int32_t nchr, int32_t res, int64_t L,
                                   int64_t U, const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins,
                                   int64_t *bin_pairs, double *bin_sumdist, int64_t *totals) {
    FHC_REQUIRE(nchr >= 0 && res > 0 && nbins >= 0, FHC_E_INVALID, "fhc_host_frag_pairs: bad nchr / res / nbins");
    FHC_REQUIRE(totals != nullptr, FHC_E_INVALID, "fhc_host_frag_pairs: null totals");
    FHC_REQUIRE(nbins == 0 || (bin_lb && bin_ub && bin_pairs && bin_sumdist), FHC_E_INVALID,
                "fhc_host_frag_pairs: null bin arrays");
    int64_t noOfFrags = 0;
    for (int c = 0; c < nchr; ++c) noOfFrags += chr_n[c];  // first loop of the reference (:596-604)
    // number of distance steps per chromosome: range(0, int(maxFrag + 1), res) with maxFrag = max(mid) - res/2 (:602,:613)
    std::vector<int64_t> nsteps(nchr, 0);
    for (int c = 0; c < nchr; ++c) {
        if (chr_n[c] <= 0) continue;
        const double maxFrag = (double)chr_maxmid[c] - (double)res / 2.0;
        const int64_t stop = (int64_t)(maxFrag + 1.0);  // int(): truncation
        nsteps[c] = stop > 0 ? (stop + res - 1) / res : 0;
    }
    // the in-range window in steps: dist = k * res with L <= dist <= U (myUtils.in_range_check; -1 = unbounded)
    const int64_t kLo = L <= 0 ? 0 : (L + res - 1) / res;
    const int64_t kHi = U < 0 ? INT64_MAX : U / res;
    // The reference walks chromosome by chromosome (sorted names) and distance by distance, adding into the bin that
    // holds the distance (forward tracker, clamped to the last bin).  Every bin therefore receives its terms in the
    // order (chromosome, distance); bins are independent of each other, so each bin is filled on its own
    // in exactly that order -- the double sum is bit-identical to the sequential walk (and bins could run concurrently).
    int64_t inrange_once = 0;
    for (int c = 0; c < nchr; ++c) {
        const int64_t n = chr_n[c];
        if (n <= 0) continue;
        const int64_t k1 = nsteps[c] - 1 < kHi ? nsteps[c] - 1 : kHi;
        if (k1 >= kLo) {
            const int64_t cntk = k1 - kLo + 1;  // sum over k in [kLo, k1] of (n - k)
            inrange_once += n * cntk - (kLo + k1) * cntk / 2;
        }
    }
    // float(dist / 1e6) once per distance step instead of once per (chromosome, step): the division is the slowest
    // operation of the inner loop
    int64_t max_steps = 0;
    for (int c = 0; c < nchr; ++c) max_steps = nsteps[c] > max_steps ? nsteps[c] : max_steps;
    std::vector<double> dist_mb((size_t)max_steps);
    for (int64_t k = 0; k < max_steps; ++k) dist_mb[(size_t)k] = (double)(k * (int64_t)res) / 1000000.0;
    auto fill_bin = [&](int b) {
        // distances of bin b: lb <= k*res <= ub; the last bin also takes everything beyond its ub (tracker clamp), and a
        // distance below bin_lb[0] cannot occur (bin_lb[0] == 0)
        int64_t kb0 = (bin_lb[b] + res - 1) / res;
        int64_t kb1 = (b == nbins - 1) ? INT64_MAX : bin_ub[b] / res;
        if (kb0 < kLo) kb0 = kLo;
        if (kb1 > kHi) kb1 = kHi;
        int64_t pairs = bin_pairs[b];
        double sumdist = bin_sumdist[b];
        for (int c = 0; c < nchr; ++c) {
            const int64_t n = chr_n[c];
            if (n <= 0) continue;
            const int64_t k1 = nsteps[c] - 1 < kb1 ? nsteps[c] - 1 : kb1;
            for (int64_t k = kb0; k <= k1; ++k) {
                const int64_t npairs = n - k;  // may go negative when loci are unmappable
                pairs += npairs;                                           // [7] and [1] (:639-640)
                sumdist += dist_mb[(size_t)k] * (double)npairs;            // :641  float(dist / 1e6) * npairs
            }
        }
        bin_pairs[b] = pairs;
        bin_sumdist[b] = sumdist;
    };
    // bins are independent and each keeps the reference's (chromosome, distance) order, so they can be filled
    // concurrently without changing a bit; the wide long-distance bins hold most of the steps, hence the work-balanced
    // split.  Small tables stay on the calling thread.
    std::vector<int64_t> work((size_t)(nbins > 0 ? nbins : 0), 0);
    int64_t total_work = 0;
    for (int b = 0; b < nbins; ++b) {
        int64_t kb0 = (bin_lb[b] + res - 1) / res;
        int64_t kb1 = (b == nbins - 1) ? INT64_MAX : bin_ub[b] / res;
        if (kb0 < kLo) kb0 = kLo;
        if (kb1 > kHi) kb1 = kHi;
        for (int c = 0; c < nchr; ++c) {
            if (chr_n[c] <= 0) continue;
            const int64_t k1 = nsteps[c] - 1 < kb1 ? nsteps[c] - 1 : kb1;
            if (k1 >= kb0) work[(size_t)b] += k1 - kb0 + 1;
        }
        total_work += work[(size_t)b];
    }
    // 4 threads by default (B200 host, 5 kb whole genome: 0.22 ms against 0.41 ms on one thread; 8 threads gain nothing
    // more); FHC_HOST_THREADS overrides
    int nthreads = 4;
    if (const char *e = getenv("FHC_HOST_THREADS")) nthreads = atoi(e) > 0 ? (atoi(e) > 64 ? 64 : atoi(e)) : 1;
    if (total_work < 200000 || nbins < 2) nthreads = 1;
    if (nthreads <= 1) {
        for (int b = 0; b < nbins; ++b) fill_bin(b);
    } else {
        // dynamic hand-out, heaviest bins first
        std::vector<int> order((size_t)nbins);
        for (int b = 0; b < nbins; ++b) order[(size_t)b] = b;
        std::sort(order.begin(), order.end(), [&](int a, int b2) { return work[(size_t)a] > work[(size_t)b2]; });
        std::atomic<int> next{0};
        auto worker = [&]() {
            for (;;) {
                const int i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= nbins) break;
                fill_bin(order[(size_t)i]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &t : pool) t.join();
    }
    int64_t interpairs2 = 0, intraall2 = 0;
    for (int c = 0; c < nchr; ++c) {
        const int64_t n = chr_n[c];
        if (n <= 0) continue;
        interpairs2 += n * (noOfFrags - n);  // :645
        intraall2 += n * (n + 1);            // :647 (x2)
    }
    // :618 adds npairs once per in-range distance, :642 once more when there are bins (the x2 of SURVEY F4)
    totals[0] = nbins > 0 ? 2 * inrange_once : inrange_once;
    totals[1] = intraall2;
    totals[2] = interpairs2;
    totals[3] = noOfFrags;
    return FHC_OK;
}

// dst[i] = v with `nthreads` threads and streaming (non-temporal) stores (host).  The end-to-end call fills its pinned q
// array with 1.0 while the GPU works; torch's own fill runs on one thread under torchrun (OMP_NUM_THREADS=1), 0.1 s per GB.
extern "C" int fhc_host_fill_f64(double *dst, int64_t n, double v, int32_t nthreads) {
    FHC_REQUIRE(n >= 0 && (n == 0 || dst != nullptr), FHC_E_INVALID, "fhc_host_fill_f64: null array or n < 0");
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (n < (1 << 20)) nthreads = 1;
    auto fill = [&](int64_t lo, int64_t hi) {
        int64_t i = lo;
        while (i < hi && (reinterpret_cast<uintptr_t>(dst + i) & 15u)) dst[i++] = v;
        const __m128d vv = _mm_set1_pd(v);
        for (; i + 8 <= hi; i += 8) {
            _mm_stream_pd(dst + i, vv);
            _mm_stream_pd(dst + i + 2, vv);
            _mm_stream_pd(dst + i + 4, vv);
            _mm_stream_pd(dst + i + 6, vv);
        }
        for (; i < hi; ++i) dst[i] = v;
        _mm_sfence();
    };
    if (nthreads == 1) {
        fill(0, n);
        return FHC_OK;
    }
    std::vector<std::thread> pool;
    const int64_t per = ((n + nthreads - 1) / nthreads + 7) & ~(int64_t)7;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(fill, per * t < n ? per * t : n, per * (t + 1) < n ? per * (t + 1) : n);
    fill(0, per < n ? per : n);
    for (auto &th : pool) th.join();
    return FHC_OK;
}

// generate_FragPairs, restriction-fragment branch (-r 0), fithic/fithic.py:691-778.  mids: the ascending mid points of the
// mappable fragments, chromosome after chromosome in the reference's sorted-name order (chr_off[nchr + 1]).  Every pair
// (x < y) of one chromosome with L <= mid_y - mid_x <= U counts once in bin_pairs1 (`[1] += 1`, :734), adds
// npairs = templen - d to bin_pairs7 (`[7]`, :733; d = in-range partners of x seen so far, :723-724 -- the reference's own
// quirk) and float(dist / 1e6) * npairs to bin_sumdist (:735), in the reference's order (x ascending, y ascending), so the
// double sums are bit identical.  The reference walks all y > x and skips the out-of-range ones; the mids are sorted, so
// the partners in range are one window per x, found by binary search.  bin_pairs1 / bin_pairs7 carry the pass >= 2 outlier
// decrements on entry.  totals: [0] possibleIntraInRangeCount, [1] possibleIntraAllCount, [2] sum n (noOfFrags - n)
// (= 2 possibleInterAllCount), [3] noOfFrags, [4] maxPossibleGenomicDist.
extern "C" int fhc_host_frag_pairs_varsize(const int64_t *mids, const int64_t *chr_off, int32_t nchr, int64_t L, int64_t U,
                                           const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins, int64_t *bin_pairs1,
                                           int64_t *bin_pairs7, double *bin_sumdist, int64_t *totals) {
    FHC_REQUIRE(nchr >= 0 && nbins >= 0 && totals != nullptr, FHC_E_INVALID, "fhc_host_frag_pairs_varsize: bad nchr / nbins / totals");
    FHC_REQUIRE(nchr == 0 || (mids && chr_off), FHC_E_INVALID, "fhc_host_frag_pairs_varsize: null fragment arrays");
    FHC_REQUIRE(nbins == 0 || (bin_lb && bin_ub && bin_pairs1 && bin_pairs7 && bin_sumdist), FHC_E_INVALID,
                "fhc_host_frag_pairs_varsize: null bin arrays");
    const int64_t noOfFrags = nchr > 0 ? chr_off[nchr] - chr_off[0] : 0;
    int64_t inrange = 0, intra_all = 0, inter2 = 0, maxdist = 0;
    for (int c = 0; c < nchr; ++c) {
        const int64_t *f = mids + chr_off[c];
        const int64_t n = chr_off[c + 1] - chr_off[c];
        if (n <= 0) continue;
        for (int64_t i = 1; i < n; ++i)
            FHC_REQUIRE(f[i] >= f[i - 1], FHC_E_INVALID, "fhc_host_frag_pairs_varsize: mid points of chromosome %d are not sorted", c);
        inter2 += (noOfFrags - n) * n;  // :701
        for (int64_t x = 0; x < n; ++x) {
            // window of partners in range (in_range_check, myUtils.py: L == -1 / U == -1 mean unbounded)
            const int64_t *lo = L < 0 ? f + x + 1 : std::lower_bound(f + x + 1, f + n, f[x] + L);
            const int64_t *hi = U < 0 ? f + n : std::upper_bound(f + x + 1, f + n, f[x] + U);
            int tr = 0;
            int64_t d = 0;
            for (const int64_t *y = lo; y < hi; ++y) {
                const int64_t di = *y - f[x];
                inrange += 1;                       // :708
                if (di > maxdist) maxdist = di;     // :711
                const int64_t npairs = n - d;       // :713
                d += 1;
                if (nbins > 0) {
                    while (!(bin_lb[tr] <= di && di <= bin_ub[tr])) {  // :719-731 forward tracker, clamped to the last bin
                        tr += 1;
                        if (tr >= nbins) {
                            tr -= 1;
                            break;
                        }
                    }
                    bin_pairs7[tr] += npairs;
                    bin_pairs1[tr] += 1;
                    bin_sumdist[tr] += ((double)di / 1000000.0) * (double)npairs;
                    intra_all += 1;                 // :736
                }
            }
        }
    }
    totals[0] = inrange;
    totals[1] = intra_all;
    totals[2] = inter2;
    totals[3] = noOfFrags;
    totals[4] = maxdist;
    return FHC_OK;
}

// ---- lbeta table with the C library's log ---------------------------------------------------------------------------------
// tab[c] = cephes lbeta(c, N - c + 1) like fhc_lbeta_table, but with log() from the C library of THIS machine -- the function
// scipy's cephes calls -- instead of the correctly rounded log of the device kernel.  The two tables differ in the rare
// entries where the library's log is not correctly rounded (about one argument in 15,000 at some magnitudes), by one ulp of
// lgam(N) ~ 4e-6 ... 8e-6 in the p-value; with this table K3 follows scipy there too.
namespace fhc {
namespace libm {
inline double c_library_log(double x) { return ::log(x); }
#define FHC_LOG_FN c_library_log
#include "cephes_lbeta.inc"
#undef FHC_LOG_FN
}  // namespace libm
}  // namespace fhc

extern "C" int fhc_host_lbeta_table(int64_t N, double *tab, int64_t ntab, int32_t nthreads) {
    FHC_REQUIRE(tab != nullptr && ntab > 0, FHC_E_INVALID, "fhc_host_lbeta_table: null table or ntab <= 0");
    FHC_REQUIRE(N >= 0 && N < (1ll << 31), FHC_E_RANGE,
                "fhc_host_lbeta_table: N = %lld does not fit the int32 that scipy.special.bdtrc truncates n to", (long long)N);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (ntab < 2048) nthreads = 1;
    auto run = [&](int64_t lo, int64_t hi) {
        for (int64_t c = lo; c < hi; ++c) {
            tab[c] = (c >= 1 && c <= N) ? fhc::libm::lbeta_cephes((double)c, (double)(N - c + 1)) : NAN;  // as lbeta_table_kernel
        }
    };
    std::vector<std::thread> pool;
    const int64_t per = (ntab + nthreads - 1) / nthreads;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(run, per * t < ntab ? per * t : ntab, per * (t + 1) < ntab ? per * (t + 1) : ntab);
    run(0, per < ntab ? per : ntab);
    for (auto &th : pool) th.join();
    return FHC_OK;
}

// ---- smoothing-spline fit ---------------------------------------------------------------------------------------------
// UnivariateSpline(x, y, s=s) of fit_Spline (fithic/fithic.py:951; scipy FITPACK curfit, see fitpack_host.cuh): cubic, unit
// weights, boundary knots at x[0] and x[m-1].  t, c [host]: room for m + 4 doubles each; the first *n_out are the knots and
// the zero-padded b-spline coefficients (`ius._eval_args`).  *ier_out is FITPACK's ier of the last run (<= 0: fine, 1..3:
// the warnings scipy prints), *calls_out 1 or 2 (2: the first run hit its storage limit, as _reset_nest handles it).
extern "C" int fhc_host_curfit(const double *x, const double *y, int32_t m, double s, double *t, double *c, int32_t *n_out,
                               double *fp_out, int32_t *ier_out, int32_t *calls_out) {
    FHC_REQUIRE(x && y && t && c && n_out && ier_out, FHC_E_INVALID, "fhc_host_curfit: null pointer");
    FHC_REQUIRE(m > 3, FHC_E_INVALID, "fhc_host_curfit: a cubic spline needs more than 3 points (m = %d)", m);
    FHC_REQUIRE(s >= 0.0, FHC_E_INVALID, "fhc_host_curfit: s must be >= 0");
    for (int i = 1; i < m; ++i)
        FHC_REQUIRE(s > 0.0 ? x[i] >= x[i - 1] : x[i] > x[i - 1], FHC_E_INVALID, "fhc_host_curfit: x must be increasing");
    int n = 0, calls = 0;
    double fp = 0.0;
    const int ier = fhc::fitpack::univariate_spline(x, y, m, 3, s, t, c, &n, &fp, &calls);
    *n_out = n;
    *ier_out = ier;
    if (fp_out) *fp_out = fp;
    if (calls_out) *calls_out = calls;
    return FHC_OK;
}
