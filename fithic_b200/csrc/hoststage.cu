// The host stage between K1 and K3 as ONE call: distance histogram in, lookup table out.
//
//   read_Interactions' dictionary (observed distances)   fithic/fithic.py:428-441
//   makeBinsFromInteractions                             fithic/fithic.py:463-553
//   generate_FragPairs, fixed-size branch                fithic/fithic.py:596-689
//   calculateProbabilities                               fithic/fithic.py:869-908
//   fit_Spline: sort, UnivariateSpline, ius(splineX), IsotonicRegression, the clamp + bisect lookup of the per-line loop
//                                                        fithic/fithic.py:936-968, :1066-1068
//   the lbeta(count, N - count + 1) table of scipy.special.bdtrc (call sites :1070, :1101)
//
// Every rank of a multi-GPU run repeats this stage with the GPU idle, so it is the floor of the strong-scaling curve
// (round 1: 1.5 ms of numpy / scipy / ctypes calls per pass against 1.4 ms of kernels at 8 GPUs).  Here it is plain C++
// on the arrays the device->host copy of K1's buffer delivers, with a small pool of spinning worker threads for the three
// parts that have independent pieces (possible-pair sums per bin, spline evaluation per point, lbeta per count).  All
// arithmetic keeps the reference's order of operations (bit exact bins, x, y, knots, table); see host_bins.cu,
// fitpack_host.cuh and spline.cu for the pieces this file strings together.
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <math.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "cephes_dev.cuh"
#include "common.cuh"
#include "fitpack_host.cuh"
#include "hostpool.cuh"
#include "shmcomm.cuh"

namespace fhc {

namespace libm2 {
inline double c_library_log(double x) { return ::log(x); }
#define FHC_LOG_FN c_library_log
#include "cephes_lbeta.inc"
#undef FHC_LOG_FN
}  // namespace libm2

static inline double wall_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

// FITPACK splev for k = 3 at one abscissa (the host twin of splev3 in spline.cu: same operations, no contraction).
// `l` is a hint: the knot interval of the previous (smaller) abscissa.
static inline double splev3_host(const double *t, const double *c, int nt, double x, int &l) {
    const int lmax = nt - 5;
    if (l < 3) l = 3;
    while (l < lmax && t[l + 1] <= x) ++l;  // largest l in [3, nt - 5] with t[l] <= x
    double h[4] = {1.0, 0.0, 0.0, 0.0}, hh[3];
    for (int j = 1; j <= 3; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 0; i < j; ++i) {
            const int li = l + 1 + i, lj = li - j;
            if (t[li] == t[lj]) {
                h[i + 1] = 0.0;
            } else {
                const double f = hh[i] / (t[li] - t[lj]);
                h[i] = h[i] + f * (t[li] - x);
                h[i + 1] = f * (x - t[lj]);
            }
        }
    }
    double sp = 0.0;
    for (int j = 0; j < 4; ++j) sp = sp + c[l - 3 + j] * h[j];
    return sp;
}

#if defined(__SSE2__)
// Two abscissae of the same knot interval l at once, the recurrence of fpbspl written out for k = 3 (lane by lane the
// operations of splev3_host; t[l] < t[l + 1], so none of the knot differences below is zero and the "coincident knots"
// branch of fpbspl is never taken).
struct SplevInterval {
    __m128d tm2, tm1, t0, t1, t2, t3;        // t[l - 2] ... t[l + 3]
    __m128d d10, d1m1, d20, d1m2, d2m1, d30;  // t[l+1]-t[l], t[l+1]-t[l-1], t[l+2]-t[l], t[l+1]-t[l-2], t[l+2]-t[l-1], t[l+3]-t[l]
    __m128d c0, c1, c2, c3;
    void set(const double *t, const double *c, int l) {
        tm2 = _mm_set1_pd(t[l - 2]); tm1 = _mm_set1_pd(t[l - 1]); t0 = _mm_set1_pd(t[l]);
        t1 = _mm_set1_pd(t[l + 1]); t2 = _mm_set1_pd(t[l + 2]); t3 = _mm_set1_pd(t[l + 3]);
        d10 = _mm_set1_pd(t[l + 1] - t[l]); d1m1 = _mm_set1_pd(t[l + 1] - t[l - 1]); d20 = _mm_set1_pd(t[l + 2] - t[l]);
        d1m2 = _mm_set1_pd(t[l + 1] - t[l - 2]); d2m1 = _mm_set1_pd(t[l + 2] - t[l - 1]); d30 = _mm_set1_pd(t[l + 3] - t[l]);
        c0 = _mm_set1_pd(c[l - 3]); c1 = _mm_set1_pd(c[l - 2]); c2 = _mm_set1_pd(c[l - 1]); c3 = _mm_set1_pd(c[l]);
    }
    __m128d eval(__m128d x) const {
        const __m128d zero = _mm_setzero_pd(), one = _mm_set1_pd(1.0);
        // j = 1
        __m128d f = _mm_div_pd(one, d10);
        __m128d h0 = _mm_add_pd(zero, _mm_mul_pd(f, _mm_sub_pd(t1, x)));
        __m128d h1 = _mm_mul_pd(f, _mm_sub_pd(x, t0));
        // j = 2
        __m128d g0 = h0, g1 = h1;
        f = _mm_div_pd(g0, d1m1);
        h0 = _mm_add_pd(zero, _mm_mul_pd(f, _mm_sub_pd(t1, x)));
        h1 = _mm_mul_pd(f, _mm_sub_pd(x, tm1));
        f = _mm_div_pd(g1, d20);
        h1 = _mm_add_pd(h1, _mm_mul_pd(f, _mm_sub_pd(t2, x)));
        __m128d h2 = _mm_mul_pd(f, _mm_sub_pd(x, t0));
        // j = 3
        g0 = h0; g1 = h1;
        const __m128d g2 = h2;
        f = _mm_div_pd(g0, d1m2);
        h0 = _mm_add_pd(zero, _mm_mul_pd(f, _mm_sub_pd(t1, x)));
        h1 = _mm_mul_pd(f, _mm_sub_pd(x, tm2));
        f = _mm_div_pd(g1, d2m1);
        h1 = _mm_add_pd(h1, _mm_mul_pd(f, _mm_sub_pd(t2, x)));
        h2 = _mm_mul_pd(f, _mm_sub_pd(x, tm1));
        f = _mm_div_pd(g2, d30);
        h2 = _mm_add_pd(h2, _mm_mul_pd(f, _mm_sub_pd(t3, x)));
        const __m128d h3 = _mm_mul_pd(f, _mm_sub_pd(x, t0));
        __m128d sp = _mm_add_pd(zero, _mm_mul_pd(c0, h0));
        sp = _mm_add_pd(sp, _mm_mul_pd(c1, h1));
        sp = _mm_add_pd(sp, _mm_mul_pd(c2, h2));
        return _mm_add_pd(sp, _mm_mul_pd(c3, h3));
    }
};
#endif

// possible-pair terms of one bin in the reference's order (chromosome, distance): a generator that hands out blocks
struct BinTerms {
    const int64_t *chr_n, *nsteps;
    const double *dist_mb;
    int nchr;
    int64_t kb0, kb1;
    int c;
    int64_t k;
    void start(int64_t kb0_, int64_t kb1_) {
        kb0 = kb0_;
        kb1 = kb1_;
        c = 0;
        k = kb0_;
    }
    int fill(double *out, int cap) {
        int n = 0;
        while (c < nchr && n < cap) {
            const int64_t nn = chr_n[c];
            const int64_t k1 = nsteps[c] - 1 < kb1 ? nsteps[c] - 1 : kb1;
            if (nn <= 0 || k > k1) {
                ++c;
                k = kb0;
                continue;
            }
            int64_t take = k1 - k + 1;
            if (take > cap - n) take = cap - n;
            const double *dm = dist_mb + k;
            const int64_t base = nn - k;
            // float(dist / 1e6) * npairs (:641); npairs = base - i as a double without a conversion per element (exact:
            // integers far below 2^53)
            int64_t i = 0;
#if defined(__SSE2__)
            __m128d v = _mm_set_pd((double)(base - 1), (double)base);
            const __m128d two = _mm_set1_pd(2.0);
            for (; i + 1 < take; i += 2) {
                _mm_storeu_pd(out + n + i, _mm_mul_pd(_mm_loadu_pd(dm + i), v));
                v = _mm_sub_pd(v, two);
            }
#endif
            for (; i < take; ++i) out[n + i] = dm[i] * (double)(base - i);
            n += (int)take;
            k += take;
        }
        return n;
    }
};

}  // namespace fhc

// One spline pass's host stage.  phases: 1 = observed distances + bins, 2 = possible pairs + probabilities (+ lbeta
// tables), 4 = spline fit + table + lookup table.  See include/fithic_b200.h for the struct.
extern "C" int fhc_host_stage(fhc_stage_io *io, int32_t phases) {
    using namespace fhc;
    FHC_REQUIRE(io != nullptr, FHC_E_INVALID, "fhc_host_stage: null io");
    const int64_t D = io->D;
    const int res = io->grid;
    const int noOfBins = io->noOfBins;
    FHC_REQUIRE(D > 0 && res > 0 && noOfBins > 0, FHC_E_INVALID, "fhc_host_stage: need D > 0, grid > 0, noOfBins > 0");
    int nworkers = io->nthreads - 1;
    if (nworkers < 0) nworkers = 0;
    if (nworkers > 63) nworkers = 63;
    HostPool &pool = HostPool::get();
    const double t_begin = wall_ms();
    io->status = 0;

    // ---- phase 1: observed distances and equal-occupancy bins ----
    if (phases & 1) {
        FHC_REQUIRE(io->k1buf && io->dists && io->sums && io->bin_lb && io->bin_ub && io->bin_sumcc, FHC_E_INVALID,
                    "fhc_host_stage: null buffer (phase 1)");
        const uint64_t *hist = io->k1buf;
        const uint64_t *scal = io->k1buf + D;
        const uint32_t *present = reinterpret_cast<const uint32_t *>(io->k1buf + D + FHC_N_SCALARS + io->n_rank_slots);
        const bool any_present = scal[FHC_S_NONPOS_LINES] != 0;
        // The observed distances (a distance whose counts sum to zero still counts as seen, :434-436 -- rare, the bitmap is
        // only read when K1 met such a line), compacted chunk by chunk on the pool together with the running sum of their
        // counts; then the equal-occupancy bins (fhc_host_make_bins' logic, fithic/fithic.py:463-553) by bisection on that
        // running sum: a bin closes at the first distance where the counts gathered since the last closure reach
        // `desired` -- `(double)cc >= desired` (:481) implies `(double)(acc + cc) >= desired` (:486), counts being >= 0, so
        // one monotone predicate decides.  All integers: the result does not depend on the chunking.
        const int64_t N = (int64_t)scal[FHC_S_INTRA_INRANGE_SUM];
        constexpr int64_t kChunk = 4096;
        const int nchunks = (int)((D + kChunk - 1) / kChunk);
        static thread_local std::vector<int64_t> chunk_cnt, chunk_sum, cum_buf;
        chunk_cnt.assign((size_t)nchunks + 1, 0);
        chunk_sum.assign((size_t)nchunks + 1, 0);
        if ((int64_t)cum_buf.size() < D) cum_buf.resize((size_t)D);
        int64_t *ccnt = chunk_cnt.data(), *csum = chunk_sum.data(), *cum = cum_buf.data();
        auto is_seen = [&](int64_t k, uint64_t v) { return v != 0 || (any_present && ((present[k >> 5] >> (k & 31)) & 1u)); };
        pool.parallel_for(nchunks, nworkers, [&](int j) {
            const int64_t k0 = (int64_t)j * kChunk, k1 = k0 + kChunk < D ? k0 + kChunk : D;
            int64_t c = 0, sum = 0;
            for (int64_t k = k0; k < k1; ++k) {
                const uint64_t v = hist[k];
                c += is_seen(k, v) ? 1 : 0;
                sum += (int64_t)v;
            }
            ccnt[j + 1] = c;
            csum[j + 1] = sum;
        });
        for (int j = 0; j < nchunks; ++j) {
            ccnt[j + 1] += ccnt[j];
            csum[j + 1] += csum[j];
        }
        const int64_t m = ccnt[nchunks];
        int64_t *dists = io->dists, *sums = io->sums;
        pool.parallel_for(nchunks, nworkers, [&](int j) {
            const int64_t k0 = (int64_t)j * kChunk, k1 = k0 + kChunk < D ? k0 + kChunk : D;
            int64_t at = ccnt[j], run = csum[j];
            int64_t sink[3];  // where the stores of an unobserved distance go (never into another chunk's share)
            for (int64_t k = k0; k < k1; ++k) {
                const uint64_t v = hist[k];
                const bool sn = is_seen(k, v);
                run += (int64_t)v;
                *(sn ? dists + at : sink) = k * (int64_t)res;
                *(sn ? sums + at : sink + 1) = (int64_t)v;
                *(sn ? cum + at : sink + 2) = run;
                at += sn ? 1 : 0;
            }
        });
        io->nseen = m;
        double desired = (double)N / (double)noOfBins;  // :476
        int64_t closed_at = 0, prev_ub = -1, i0 = 0;     // counts up to the last closure; its distance; first open entry
        int nb = 0;
        while (i0 < m) {
            int64_t lo2 = i0, hi2 = m;  // first i >= i0 with (double)(cum[i] - closed_at) >= desired
            while (lo2 < hi2) {
                const int64_t mid = (lo2 + hi2) >> 1;
                if ((double)(cum[mid] - closed_at) >= desired) hi2 = mid; else lo2 = mid + 1;
            }
            if (lo2 >= m) break;  // distances after the last closed bin are dropped, as in the reference
            if (nb >= noOfBins) {
                fhc::set_error("fhc_host_stage: more than noOfBins=%d bins closed", noOfBins);
                return FHC_E_RANGE;
            }
            io->bin_lb[nb] = nb == 0 ? 0 : prev_ub + 1;  // :518-521
            io->bin_ub[nb] = dists[lo2];
            io->bin_sumcc[nb] = cum[lo2] - closed_at;
            prev_ub = dists[lo2];
            closed_at = cum[lo2];
            nb += 1;
            if (nb < noOfBins) desired = 1.0 * (double)(N - closed_at) / (double)(noOfBins - nb);  // :500-502
            i0 = lo2 + 1;
        }
        io->nb = nb;
        io->timings[0] = wall_ms() - t_begin;
    }

    // ---- phase 2: possible pairs per bin, probabilities; lbeta tables ride along on the pool ----
    if (phases & (2 | 8)) {  // 8: the lbeta tables alone (after the caller enlarged them)
        const double t0 = wall_ms();
        const bool do_bins = (phases & 2) != 0;
        const int nb = do_bins ? io->nb : 0;
        const int nchr = io->nchr;
        const int64_t L = io->L, U = io->U;
        const uint64_t *scal = io->k1buf + D;
        const int64_t N = (int64_t)scal[FHC_S_INTRA_INRANGE_SUM];
        FHC_REQUIRE(nchr == 0 || (io->chr_n && io->chr_maxmid), FHC_E_INVALID, "fhc_host_stage: null fragment arrays");
        FHC_REQUIRE(!do_bins || (io->bin_pairs && io->bin_sumdist && io->x_bins && io->y_bins), FHC_E_INVALID,
                    "fhc_host_stage: null buffer (phase 2)");
        int64_t noOfFrags = 0;
        for (int c = 0; c < nchr; ++c) noOfFrags += io->chr_n[c];
        std::vector<int64_t> nsteps((size_t)nchr, 0);
        int64_t max_steps = 0;
        for (int c = 0; c < nchr; ++c) {
            if (io->chr_n[c] <= 0) continue;
            const double maxFrag = (double)io->chr_maxmid[c] - (double)res / 2.0;
            const int64_t stop = (int64_t)(maxFrag + 1.0);
            nsteps[(size_t)c] = stop > 0 ? (stop + res - 1) / res : 0;
            if (nsteps[(size_t)c] > max_steps) max_steps = nsteps[(size_t)c];
        }
        const int64_t kLo = L <= 0 ? 0 : (L + res - 1) / res;
        const int64_t kHi = U < 0 ? INT64_MAX : U / res;
        int64_t inrange_once = 0, interpairs2 = 0, intraall2 = 0;
        for (int c = 0; c < nchr; ++c) {
            const int64_t n = io->chr_n[c];
            if (n <= 0) continue;
            const int64_t k1 = nsteps[(size_t)c] - 1 < kHi ? nsteps[(size_t)c] - 1 : kHi;
            if (k1 >= kLo) {
                const int64_t cntk = k1 - kLo + 1;
                inrange_once += n * cntk - (kLo + k1) * cntk / 2;
            }
            interpairs2 += n * (noOfFrags - n);
            intraall2 += n * (n + 1);
        }
        if (do_bins) {
            io->totals[0] = nb > 0 ? 2 * inrange_once : inrange_once;  // the x2 of :618 + :642
            io->totals[1] = intraall2;
            io->totals[2] = interpairs2;
            io->totals[3] = noOfFrags;
        }

        // float(dist / 1e6) per distance step (:641), kept between calls: 50 k divisions are 0.07 ms
        static thread_local std::vector<double> dist_mb;
        static thread_local int dist_mb_res = 0;
        if ((int64_t)dist_mb.size() < max_steps || dist_mb_res != res) {
            dist_mb.resize((size_t)max_steps);
            for (int64_t k = 0; k < max_steps; ++k) dist_mb[(size_t)k] = (double)(k * (int64_t)res) / 1000000.0;
            dist_mb_res = res;
        }
        // per bin: the window of distance steps, the integer pair count in closed form, the work of its double sum
        struct BinJob {
            int b;
            int64_t kb0, kb1, work;
        };
        std::vector<BinJob> jobs((size_t)nb);
        for (int b = 0; b < nb; ++b) {
            int64_t kb0 = (io->bin_lb[b] + res - 1) / res;
            int64_t kb1 = (b == nb - 1) ? INT64_MAX : io->bin_ub[b] / res;
            if (kb0 < kLo) kb0 = kLo;
            if (kb1 > kHi) kb1 = kHi;
            int64_t pairs = io->dec ? -io->dec[b] : 0;  // pass >= 2: the outlier decrements hit [1] and [7] (:544-545)
            int64_t work = 0;
            for (int c = 0; c < nchr; ++c) {
                const int64_t n = io->chr_n[c];
                if (n <= 0) continue;
                const int64_t k1 = nsteps[(size_t)c] - 1 < kb1 ? nsteps[(size_t)c] - 1 : kb1;
                if (k1 >= kb0) {
                    const int64_t cntk = k1 - kb0 + 1;
                    pairs += n * cntk - (kb0 + k1) * cntk / 2;
                    work += cntk;
                }
            }
            io->bin_pairs[b] = pairs;
            jobs[(size_t)b] = {b, kb0, kb1, work};
        }
        std::sort(jobs.begin(), jobs.end(), [](const BinJob &a, const BinJob &b2) { return a.work > b2.work; });
        // Several GPUs: every rank would repeat the same sums, so each takes a share of the bins instead (heaviest first
        // onto the least loaded rank: the same assignment on every rank) and leaves 0.0 in the others; the caller adds the
        // arrays of all ranks up (x + 0.0 = x: the sums stay the bits one rank computed, in the reference's order).
        if (io->pairs_world > 1) {
            std::vector<int64_t> load((size_t)io->pairs_world, 0);
            std::vector<BinJob> mine;
            for (const BinJob &jb : jobs) {
                int r = 0;
                for (int q = 1; q < io->pairs_world; ++q)
                    if (load[(size_t)q] < load[(size_t)r]) r = q;
                load[(size_t)r] += jb.work;
                if (r == io->pairs_rank)
                    mine.push_back(jb);
                else
                    io->bin_sumdist[jb.b] = 0.0;
            }
            jobs.swap(mine);
        }
        const int nb_jobs = (int)jobs.size();
        // the double sums: a bin is one chain of dependent additions (the order is the reference's); eight bins of similar
        // length are summed side by side so that the adder pipelines stay full
        constexpr int kChains = 8;  // two FP adders x four cycles of latency
        const int ngroups = (nb_jobs + kChains - 1) / kChains;
        // lbeta tables
        int64_t lb_N[2] = {0, 0}, lb_n[2] = {0, 0};
        int64_t max_count = (int64_t)scal[FHC_S_MAX_COUNT];
        for (int r = 0; r < io->n_rank_slots; ++r)  // multi-GPU: every rank's maximum arrives in its own slot
            if ((int64_t)scal[FHC_N_SCALARS + r] > max_count) max_count = (int64_t)scal[FHC_N_SCALARS + r];
        for (int w = 0; w < 2; ++w) {
            if (io->lbeta_tab[w] == nullptr) continue;
            const int64_t Nw = w == 0 ? N : (int64_t)scal[FHC_S_INTER_ALL_SUM];
            FHC_REQUIRE(Nw >= 0 && Nw < (1ll << 31), FHC_E_RANGE,
                        "N = %lld does not fit the int32 that scipy.special.bdtrc truncates n to", (long long)Nw);
            int64_t cap = Nw < ((1ll << 22) - 1) ? Nw : ((1ll << 22) - 1);
            int64_t ntab = (max_count > 1 ? max_count : 1);
            if (ntab > cap) ntab = cap;
            ntab += 1;
            if (ntab > io->lbeta_cap[w]) {
                io->status = 3;  // the caller's table is too small: io->lbeta_ntab says what is needed
                io->lbeta_ntab[w] = ntab;
                ntab = 0;
            } else {
                io->lbeta_ntab[w] = ntab;
            }
            lb_N[w] = Nw;
            lb_n[w] = ntab;
        }
        const int kLbChunk = 256;
        const int lbjobs0 = (int)((lb_n[0] + kLbChunk - 1) / kLbChunk), lbjobs1 = (int)((lb_n[1] + kLbChunk - 1) / kLbChunk);
        const int64_t *chr_n = io->chr_n;
        const double *dist_mb_p = dist_mb.data();  // (a thread_local named inside the lambda would be the worker's own)
        auto run_job = [&](int j) {
            if (j < ngroups) {
                const int g0 = j * kChains;
                const int ng = nb_jobs - g0 < kChains ? nb_jobs - g0 : kChains;
                constexpr int kBlk = 256;
                double buf[kChains][kBlk];
                int len[kChains], pos[kChains];
                double s[kChains];
                bool live[kChains];
                BinTerms gen[kChains];
                for (int q = 0; q < kChains; ++q) {
                    len[q] = pos[q] = 0;
                    s[q] = 0.0;
                    live[q] = q < ng;
                    if (q < ng) {
                        gen[q].chr_n = chr_n;
                        gen[q].nsteps = nsteps.data();
                        gen[q].dist_mb = dist_mb_p;
                        gen[q].nchr = nchr;
                        gen[q].start(jobs[(size_t)(g0 + q)].kb0, jobs[(size_t)(g0 + q)].kb1);
                    }
                }
                for (;;) {
                    int mn = INT32_MAX, nlive = 0;
                    for (int q = 0; q < ng; ++q) {
                        if (!live[q]) continue;
                        if (pos[q] == len[q]) {
                            len[q] = gen[q].fill(buf[q], kBlk);
                            pos[q] = 0;
                            if (len[q] == 0) {
                                live[q] = false;
                                continue;
                            }
                        }
                        ++nlive;
                        if (len[q] - pos[q] < mn) mn = len[q] - pos[q];
                    }
                    if (nlive == 0) break;
                    if (nlive == kChains) {
                        const double *b0 = buf[0] + pos[0], *b1 = buf[1] + pos[1], *b2 = buf[2] + pos[2], *b3 = buf[3] + pos[3];
                        const double *b4 = buf[4] + pos[4], *b5 = buf[5] + pos[5], *b6 = buf[6] + pos[6], *b7 = buf[7] + pos[7];
                        double s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3], s4 = s[4], s5 = s[5], s6 = s[6], s7 = s[7];
                        for (int i = 0; i < mn; ++i) {
                            s0 += b0[i]; s1 += b1[i]; s2 += b2[i]; s3 += b3[i];
                            s4 += b4[i]; s5 += b5[i]; s6 += b6[i]; s7 += b7[i];
                        }
                        s[0] = s0; s[1] = s1; s[2] = s2; s[3] = s3; s[4] = s4; s[5] = s5; s[6] = s6; s[7] = s7;
                        for (int q = 0; q < kChains; ++q) pos[q] += mn;
                    } else {
                        // the chains that are left, two at a time
                        int qa = -1;
                        for (int q = 0; q < ng; ++q) {
                            if (!live[q]) continue;
                            if (qa < 0) {
                                qa = q;
                                continue;
                            }
                            const double *ba = buf[qa] + pos[qa], *bb = buf[q] + pos[q];
                            double sa = s[qa], sb = s[q];
                            for (int i = 0; i < mn; ++i) {
                                sa += ba[i];
                                sb += bb[i];
                            }
                            s[qa] = sa;
                            s[q] = sb;
                            pos[qa] += mn;
                            pos[q] += mn;
                            qa = -1;
                        }
                        if (qa >= 0) {
                            const double *ba = buf[qa] + pos[qa];
                            double sa = s[qa];
                            for (int i = 0; i < mn; ++i) sa += ba[i];
                            s[qa] = sa;
                            pos[qa] += mn;
                        }
                    }
                }
                for (int q = 0; q < ng; ++q) io->bin_sumdist[jobs[(size_t)(g0 + q)].b] = s[q];
                return;
            }
            j -= ngroups;
            const int w = j < lbjobs0 ? 0 : 1;
            if (w == 1) j -= lbjobs0;
            const int64_t lo = (int64_t)j * kLbChunk, hi = lo + kLbChunk < lb_n[w] ? lo + kLbChunk : lb_n[w];
            double *tab = io->lbeta_tab[w];
            const int64_t Nw = lb_N[w];
            for (int64_t c = lo; c < hi; ++c)
                tab[c] = (c >= 1 && c <= Nw) ? libm2::lbeta_cephes((double)c, (double)(Nw - c + 1)) : NAN;
        };
        pool.parallel_for(ngroups + lbjobs0 + lbjobs1, nworkers, run_job);
        if (io->pairs_world > 1 && io->shm != nullptr) {
            // every entry of bin_sumdist is one rank's sum and zeros elsewhere: adding the bit patterns up is exact
            fhc::ShmComm *sc = reinterpret_cast<fhc::ShmComm *>(io->shm);
            static_assert(sizeof(double) == sizeof(unsigned long long), "bit patterns of doubles are added as integers");
            const int rc = fhc::shm_allreduce_u64(*sc, reinterpret_cast<unsigned long long *>(io->bin_sumdist), nb, 20.0);
            FHC_REQUIRE(rc == 0, FHC_E_INVALID, "fhc_host_stage: the exchange of the possible-pair sums between the ranks failed (%d)", rc);
        }
        io->timings[1] = wall_ms() - t0;
    }
    if (((phases & 2) && (io->pairs_world <= 1 || io->shm != nullptr)) || (phases & 16)) {
        // calculateProbabilities (:869-908): y = avgCC = (sumCC / pairs) / N, x = avgDist = 1e6 * (sumDist / pairs)
        const int nb = io->nb;
        const int64_t N = (int64_t)(io->k1buf + D)[FHC_S_INTRA_INRANGE_SUM];
        for (int b = 0; b < nb; ++b) {
            const int64_t pairs = io->bin_pairs[b];
            double y = 0.0, x = 0.0;
            if (pairs > 0 && N > 0) y = (1.0 * (double)io->bin_sumcc[b] / (double)pairs) / (double)N;
            if (pairs != 0) x = 1000000.0 * (io->bin_sumdist[b] / (double)pairs);
            io->x_bins[b] = x;
            io->y_bins[b] = y;
        }
    }

    // ---- phase 4: spline fit, table at the observed distances, antitonic regression, lookup table ----
    if ((phases & 4) && io->want_spline) {
        double t0 = wall_ms();
        const int nb = io->nb;
        FHC_REQUIRE(io->xs && io->ys && io->t && io->c && io->splineX && io->table && io->lut, FHC_E_INVALID,
                    "fhc_host_stage: null buffer (phase 4)");
        // sort (x, y) by x (:937-939); x must increase strictly (:940-945)
        std::vector<int> order((size_t)nb);
        for (int b = 0; b < nb; ++b) order[(size_t)b] = b;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return io->x_bins[a] < io->x_bins[b2]; });
        for (int b = 0; b < nb; ++b) {
            io->xs[b] = io->x_bins[order[(size_t)b]];
            io->ys[b] = io->y_bins[order[(size_t)b]];
        }
        for (int b = 1; b < nb; ++b) {
            if (io->xs[b] <= io->xs[b - 1]) {
                io->status = 1;  // "Distances do not decrease across bins" -> exit(2) in the reference
                io->bad_index = b;
                return FHC_OK;
            }
        }
        if (nb <= 3) {
            io->status = 4;  // scipy refuses m <= k
            return FHC_OK;
        }
        double ymin = io->ys[0];
        for (int b = 1; b < nb; ++b) ymin = io->ys[b] < ymin ? io->ys[b] : ymin;  // NaN-free: y is a ratio of counts
        const double s = ymin * ymin;  // :948
        int n = 0, calls = 0;
        double fp = 0.0;
        const int ier = fitpack::univariate_spline(io->xs, io->ys, nb, 3, s, io->t, io->c, &n, &fp, &calls);
        io->nt = n;
        io->ier = ier;
        io->calls = calls;
        io->fp = fp;
        io->timings[2] = wall_ms() - t0;
        t0 = wall_ms();
        // splineX: the observed distances inside [min(x), max(x)] (:952-960)
        const double xmin = io->xs[0], xmax = io->xs[nb - 1];
        int64_t lo = 0;
        while (lo < io->nseen && (double)io->dists[lo] < xmin) ++lo;
        int64_t hi = io->nseen;
        while (hi > lo && (double)io->dists[hi - 1] > xmax) --hi;
        const int64_t m = hi - lo;
        io->m = m;
        if (m <= 0) {
            io->status = 2;  // no observed distance falls inside the fitted range
            return FHC_OK;
        }
        memcpy(io->splineX, io->dists + lo, sizeof(int64_t) * (size_t)m);
        // ius(splineX) (:961)
        {
            const int kChunk = 2048;
            const int nj = (int)((m + kChunk - 1) / kChunk);
            const double *t = io->t, *c = io->c;
            const int64_t *sx = io->splineX;
            double *tab = io->table;
            pool.parallel_for(nj, nworkers, [&](int j) {
                const int64_t a = (int64_t)j * kChunk, b = a + kChunk < m ? a + kChunk : m;
                int l = 3;
                int64_t i = a;
#if defined(__SSE2__)
                const int lmax = n - 5;
                SplevInterval iv;
                int l_set = -1;
                for (; i + 1 < b; i += 2) {
                    const double x0 = (double)sx[i], x1 = (double)sx[i + 1];
                    while (l < lmax && t[l + 1] <= x0) ++l;
                    if ((l < lmax && t[l + 1] <= x1) || !(t[l] < t[l + 1])) {  // the pair straddles a knot: one at a time
                        tab[i] = splev3_host(t, c, n, x0, l);
                        tab[i + 1] = splev3_host(t, c, n, x1, l);
                        continue;
                    }
                    if (l != l_set) {
                        iv.set(t, c, l);
                        l_set = l;
                    }
                    _mm_storeu_pd(tab + i, iv.eval(_mm_set_pd(x1, x0)));
                }
#endif
                for (; i < b; ++i) tab[i] = splev3_host(t, c, n, (double)sx[i], l);
            });
        }
        io->timings[3] = wall_ms() - t0;
        t0 = wall_ms();
        // IsotonicRegression(increasing=False).fit_transform (:965-966)
        {
            const int rc = fhc_host_antitonic(io->table, m);
            if (rc < 0) return rc;
        }
        io->timings[4] = wall_ms() - t0;
        t0 = wall_ms();
        // lut[k] = table[min(bisect_left(splineX, clamp(k * res, xmin, xmax)), m - 1)]  (:1066-1068)
        {
            const int64_t kChunk = 8192;  // independent of the thread count: a chunk finds its first table entry by bisection
            const int nj = (int)((D + kChunk - 1) / kChunk);
            const int64_t *sx = io->splineX;
            const double *tab = io->table;
            double *lut = io->lut;
            pool.parallel_for(nj, nworkers, [&](int jb) {
                const int64_t k0 = (int64_t)jb * kChunk, k1 = k0 + kChunk < D ? k0 + kChunk : D;
                double d0 = (double)(k0 * (int64_t)res);
                d0 = d0 < xmin ? xmin : d0;
                d0 = d0 > xmax ? xmax : d0;
                int64_t lo2 = 0, hi2 = m;  // first j with sx[j] >= d0
                while (lo2 < hi2) {
                    const int64_t mid = (lo2 + hi2) >> 1;
                    if ((double)sx[mid] < d0) lo2 = mid + 1; else hi2 = mid;
                }
                int64_t j = lo2;
                for (int64_t k = k0; k < k1; ++k) {
                    double dl = (double)(k * (int64_t)res);
                    dl = dl < xmin ? xmin : dl;
                    dl = dl > xmax ? xmax : dl;
                    while (j < m && (double)sx[j] < dl) ++j;
                    lut[k] = tab[j < m ? j : m - 1];
                }
            });
        }
        io->timings[5] = wall_ms() - t0;
    }
    io->timings[7] = wall_ms() - t_begin;
    return FHC_OK;
}

extern "C" int fhc_host_pool_prewarm(int32_t nthreads) {
    int nworkers = nthreads - 1;
    if (nworkers <= 0) return FHC_OK;
    if (nworkers > 63) nworkers = 63;
    fhc::HostPool::get().ensure(nworkers);
    fhc::HostPool::get().prewarm();
    return FHC_OK;
}

// Diagnostic: njobs jobs that each spin for job_us microseconds on nthreads threads; returns the elapsed milliseconds
// (ideal: njobs * job_us / nthreads) and, in *distinct_threads, how many threads ran at least one job.
extern "C" double fhc_host_pool_selftest(int32_t nthreads, int32_t njobs, int32_t job_us, int32_t *distinct_threads) {
    using namespace fhc;
    int nworkers = nthreads - 1;
    if (nworkers < 0) nworkers = 0;
    if (nworkers > 63) nworkers = 63;
    std::mutex mu;
    std::vector<std::thread::id> ids;
    const double t0 = wall_ms();
    HostPool::get().parallel_for(njobs, nworkers, [&](int) {
        const double t = wall_ms();
        while ((wall_ms() - t) * 1e3 < (double)job_us) {
        }
        std::lock_guard<std::mutex> lk(mu);
        const auto id = std::this_thread::get_id();
        if (std::find(ids.begin(), ids.end(), id) == ids.end()) ids.push_back(id);
    });
    const double dt = wall_ms() - t0;
    if (distinct_threads) *distinct_threads = (int32_t)ids.size();
    return dt;
}

// Shared-memory exchange between the ranks of one node for the host stage (see shmcomm.cuh): rank 0 creates `name`, the
// others open it; fhc_stage_io.shm takes the handle.
extern "C" int fhc_shm_open(const char *name, int32_t rank, int32_t world, int64_t slot_bytes, void **handle_out) {
    FHC_REQUIRE(name && handle_out, FHC_E_INVALID, "fhc_shm_open: null pointer");
    fhc::ShmComm *c = new fhc::ShmComm();
    const int rc = fhc::shm_open_comm(*c, name, rank, world, slot_bytes, 60.0);
    if (rc != 0) {
        delete c;
        fhc::set_error("fhc_shm_open: cannot set up %s for rank %d of %d (%d)", name, rank, world, rc);
        return FHC_E_INVALID;
    }
    *handle_out = c;
    return FHC_OK;
}

extern "C" int fhc_shm_allreduce_u64(void *handle, uint64_t *data, int32_t n) {
    FHC_REQUIRE(handle && data && n > 0, FHC_E_INVALID, "fhc_shm_allreduce_u64: bad arguments");
    const int rc = fhc::shm_allreduce_u64(*reinterpret_cast<fhc::ShmComm *>(handle), reinterpret_cast<unsigned long long *>(data), n, 20.0);
    FHC_REQUIRE(rc == 0, FHC_E_INVALID, "fhc_shm_allreduce_u64: a rank did not arrive (%d)", rc);
    return FHC_OK;
}

extern "C" int fhc_shm_close(void *handle) {
    if (handle != nullptr) {
        fhc::ShmComm *c = reinterpret_cast<fhc::ShmComm *>(handle);
        fhc::shm_close_comm(*c);
        delete c;
    }
    return FHC_OK;
}
