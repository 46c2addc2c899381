// K2 -- spline table: FITPACK splev (de Boor) at the observed distances, antitonic regression (PAVA), dense lookup table.
//
// Replaces `splineY = ius(splineX)` and `IsotonicRegression(increasing=False).fit_transform(splineX, splineY)` of
// fit_Spline (reference fithic/fithic.py:952-966; scipy FITPACK splev/fpbspl and scipy.optimize.isotonic_regression are
// third party) and bakes the clamp + bisect of the per-contact lookup (:1066-1068) into a table indexed by distance slot.
//
// Latency-bound (m <= D <= ~50k points), one CTA:
//   1. splev: one thread per point, same operation order as fpbspl/splev, round-to-nearest without contraction -> bit
//      exact against FITPACK.
//   2. PAVA as a merge tree: 1024 threads pool adjacent violators inside their own contiguous chunk, then log2(1024)
//      rounds merge neighbouring chunks by cascading only across the shared boundary.  Blocks live in three arrays
//      (sum at block start, end-of-block at block start, start-of-block at block end), so a pool is O(1) and nothing
//      is compacted.  Pooling adjacent violators in any order gives the unique antitonic least-squares fit.
//   3. a second kernel fills lut[k] for every distance slot.
#define FHC_PROFILE_STREAM st
#include <vector>

#include "common.cuh"

namespace fhc {

constexpr int kPavaThreads = 1024;

// FITPACK splev for k = 3 at one abscissa.  t[0..nt), c[0..nt) (zero padded), x inside [t[3], t[nt-4]].
__device__ double splev3(const double *__restrict__ t, const double *__restrict__ c, int nt, double x) {
    // interval: largest l in [3, nt-5] with t[l] <= x  (splev.f: "search for knot interval t(l) <= arg < t(l+1)")
    int lo = 3, hi = nt - 5;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid] <= x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const int l = lo;
    double h[4] = {1.0, 0.0, 0.0, 0.0}, hh[3];
    for (int j = 1; j <= 3; ++j) {  // fpbspl.f
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 0; i < j; ++i) {
            const int li = l + 1 + i, lj = li - j;
            if (t[li] == t[lj]) {
                h[i + 1] = 0.0;
            } else {
                const double f = __ddiv_rn(hh[i], __dsub_rn(t[li], t[lj]));
                h[i] = __dadd_rn(h[i], __dmul_rn(f, __dsub_rn(t[li], x)));
                h[i + 1] = __dmul_rn(f, __dsub_rn(x, t[lj]));
            }
        }
    }
    double sp = 0.0;
    for (int j = 0; j < 4; ++j) sp = __dadd_rn(sp, __dmul_rn(c[l - 3 + j], h[j]));
    return sp;
}

// true when block A (sum sa over na points) followed by block B violates "non-increasing": mean(A) < mean(B)
__device__ __forceinline__ bool violates(double sa, long long na, double sb, long long nb) {
    return sa * (double)nb < sb * (double)na;
}

// A block [s, e] of pooled points is described twice, at its first and at its last index, so that a neighbour reaches
// everything it needs (sum and extent) with ONE 16-byte load: the merge cascade is a chain of dependent loads and its
// length is data dependent (a rising tail of the spline pools back over thousands of blocks), so latency per step is
// what matters.
struct __align__(16) BlockRec {
    double sum;
    int other;  // at the first index: last index of the block (-1: not a block start any more); at the last index: first
    int pad;
};

__device__ __forceinline__ BlockRec ld_rec(const BlockRec *p) {
    const int4 v = *reinterpret_cast<const int4 *>(p);
    BlockRec r;
    r.sum = __hiloint2double(v.y, v.x);
    r.other = v.z;
    r.pad = 0;
    return r;
}
__device__ __forceinline__ void st_rec(BlockRec *p, double sum, int other) {
    *reinterpret_cast<int4 *>(p) = make_int4(__double2loint(sum), __double2hiint(sum), other, 0);
}

__global__ void __launch_bounds__(kPavaThreads, 1)
spline_pava_kernel(const double *__restrict__ t, const double *__restrict__ c, int nt,
                   const long long *__restrict__ splineX, long long m, double *__restrict__ table, BlockRec *atStart,
                   BlockRec *atEnd) {
    // ---- 1 + 2a. splev, then PAVA inside each thread's chunk ----
    const long long chunk = (m + kPavaThreads - 1) / kPavaThreads;
    {
        const long long lo = (long long)threadIdx.x * chunk;
        const long long hi = lo + chunk < m ? lo + chunk : m;
        for (long long i = lo; i < hi; ++i) {
            const double y = splev3(t, c, nt, (double)splineX[i]);
            long long s = i;
            double sm = y;
            while (s > lo) {
                const BlockRec prev = ld_rec(atEnd + (s - 1));  // the block that ends just before s
                const long long ps = prev.other;
                if (!violates(prev.sum, s - ps, sm, i - s + 1)) break;
                atStart[s].other = -1;  // s stops being a block start
                sm += prev.sum;
                s = ps;
            }
            st_rec(atStart + s, sm, (int)i);
            st_rec(atEnd + i, sm, (int)s);
        }
    }
    __syncthreads();
    // ---- 2b. merge tree across chunk boundaries ----
    for (long long seg = chunk; seg < m; seg <<= 1) {
        const long long b = (2ll * threadIdx.x + 1) * seg;  // first index of the right segment
        if (b < m) {
            const long long lo = b - seg;
            const long long hi = b + seg < m ? b + seg : m;
            const BlockRec L = ld_rec(atEnd + (b - 1)), R = ld_rec(atStart + b);
            long long cs = L.other, ce = R.other;
            if (violates(L.sum, b - cs, R.sum, ce - b + 1)) {
                double sm = L.sum + R.sum;
                atStart[b].other = -1;
                bool changed = true;
                while (changed) {
                    changed = false;
                    // both neighbours are fetched before either is tested: two independent loads per round
                    BlockRec pl, nr;
                    const bool hasl = cs > lo, hasr = ce + 1 < hi;
                    if (hasl) pl = ld_rec(atEnd + (cs - 1));
                    if (hasr) nr = ld_rec(atStart + (ce + 1));
                    if (hasl && violates(pl.sum, cs - pl.other, sm, ce - cs + 1)) {
                        atStart[cs].other = -1;
                        sm += pl.sum;
                        cs = pl.other;
                        changed = true;
                    }
                    if (hasr && violates(sm, ce - cs + 1, nr.sum, nr.other - ce)) {
                        atStart[ce + 1].other = -1;
                        sm += nr.sum;
                        ce = nr.other;
                        changed = true;
                    }
                }
                st_rec(atStart + cs, sm, (int)ce);
                st_rec(atEnd + ce, sm, (int)cs);
            }
        }
        __syncthreads();
    }
    // ---- 2c. block means: propagate the last block start (max-scan over chunks), then fill ----
    __shared__ long long last_start[kPavaThreads];
    const long long lo = (long long)threadIdx.x * chunk;
    const long long hi = lo + chunk < m ? lo + chunk : m;
    long long mine = -1;
    for (long long i = lo; i < hi; ++i)
        if (atStart[i].other >= 0) mine = i;
    last_start[threadIdx.x] = mine;
    __syncthreads();
    for (int off = 1; off < kPavaThreads; off <<= 1) {  // inclusive max-scan (Hillis-Steele)
        long long v = last_start[threadIdx.x];
        if (threadIdx.x >= off) v = max(v, last_start[threadIdx.x - off]);
        __syncthreads();
        last_start[threadIdx.x] = v;
        __syncthreads();
    }
    long long cur = threadIdx.x > 0 ? last_start[threadIdx.x - 1] : -1;
    double mean = 0.0;
    if (cur >= 0) {
        const BlockRec r = ld_rec(atStart + cur);
        mean = __ddiv_rn(r.sum, (double)(r.other - cur + 1));
    }
    for (long long i = lo; i < hi; ++i) {
        const BlockRec r = ld_rec(atStart + i);
        if (r.other >= 0) mean = __ddiv_rn(r.sum, (double)(r.other - i + 1));
        table[i] = mean;
    }
}

__global__ void spline_eval_kernel(const double *__restrict__ t, const double *__restrict__ c, int nt,
                                   const long long *__restrict__ splineX, long long m, double *__restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) y[i] = splev3(t, c, nt, (double)splineX[i]);
}

__global__ void spline_lut_kernel(const long long *__restrict__ splineX, const double *__restrict__ table, long long m,
                                  double xmin, double xmax, unsigned int res, double *__restrict__ lut, long long D) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D) return;
    double dl = (double)(k * (long long)res);
    dl = fmax(dl, xmin);  // distToLookUp = max(d, min(x)); min(., max(x))   (fithic/fithic.py:1066-1067)
    dl = fmin(dl, xmax);
    long long lo = 0, hi = m;  // bisect_left: first j with splineX[j] >= dl
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((double)splineX[mid] < dl)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo > m - 1) lo = m - 1;
    lut[k] = table[lo];
}

}  // namespace fhc

extern "C" size_t fhc_spline_workspace_bytes(int64_t m) {
    if (m < 0) m = 0;
    return (size_t)m * 32 + 64;  // two 16-byte block records per point
}

extern "C" int fhc_spline_table(const double *t, const double *c, int32_t nt, const int64_t *splineX, int64_t m,
                                double xmin, double xmax, int32_t res, double *table, double *lut, int64_t D,
                                void *workspace, size_t workspace_bytes, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(t && c && splineX && table && workspace, FHC_E_INVALID, "fhc_spline_table: null pointer");
    FHC_REQUIRE(nt >= 8, FHC_E_INVALID, "fhc_spline_table: a cubic spline has at least 8 knots (got %d)", nt);
    FHC_REQUIRE(m > 0 && m < (1ll << 31), FHC_E_INVALID, "fhc_spline_table: need 0 < m < 2^31 (got %lld)", (long long)m);
    FHC_REQUIRE(res > 0 && D >= 0 && (D == 0 || lut != nullptr), FHC_E_INVALID, "fhc_spline_table: bad res / D / lut");
    FHC_REQUIRE(workspace_bytes >= fhc_spline_workspace_bytes(m), FHC_E_WORKSPACE,
                "fhc_spline_table: workspace of %zu bytes, need %zu", workspace_bytes, fhc_spline_workspace_bytes(m));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_REQUIRE(aligned16(workspace), FHC_E_INVALID, "fhc_spline_table: workspace must be 16-byte aligned");
    BlockRec *atStart = reinterpret_cast<BlockRec *>(workspace);
    BlockRec *atEnd = atStart + m;
    spline_pava_kernel<<<1, kPavaThreads, 0, st>>>(t, c, nt, reinterpret_cast<const long long *>(splineX), m, table,
                                                  atStart, atEnd);
    FHC_LAUNCH_CHECK("spline_pava_kernel");
    if (D > 0) {
        const int threads = 256;
        spline_lut_kernel<<<(unsigned int)((D + threads - 1) / threads), threads, 0, st>>>(
            reinterpret_cast<const long long *>(splineX), table, m, xmin, xmax, (unsigned int)res, lut, D);
        FHC_LAUNCH_CHECK("spline_lut_kernel");
    }
    return FHC_OK;
}

// ---- the same stages one by one ------------------------------------------------------------------------------------
// PAVA is a sequential scan whose merge cascade can run over thousands of blocks (a rising tail of the spline pools
// back over most of the curve: 5,931 dependent steps at the top of the merge tree on the 5 kb whole-genome bench
// input, 1.6 ms in spline_pava_kernel).  A CPU core does the whole scan from L1 in ~0.1 ms, so for large tables the
// engine evaluates the spline on the device (bit exact, parallel), pools on the host and builds the lookup table on
// the device again.
extern "C" int fhc_spline_eval(const double *t, const double *c, int32_t nt, const int64_t *splineX, int64_t m, double *y,
                               void *stream) {
    using namespace fhc;
    FHC_REQUIRE(t && c && splineX && y, FHC_E_INVALID, "fhc_spline_eval: null pointer");
    FHC_REQUIRE(nt >= 8 && m > 0, FHC_E_INVALID, "fhc_spline_eval: need nt >= 8 and m > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    spline_eval_kernel<<<(unsigned int)((m + 127) / 128), 128, 0, st>>>(t, c, nt, reinterpret_cast<const long long *>(splineX),
                                                                       m, y);
    FHC_LAUNCH_CHECK("spline_eval_kernel");
    return FHC_OK;
}

// IsotonicRegression(increasing=False).fit_transform with unit weights (fithic/fithic.py:965-966): pool adjacent
// violators, every point takes the mean of its block.  In place on a host array.
extern "C" int fhc_host_antitonic(double *y, int64_t m) {
    FHC_REQUIRE(m >= 0 && (m == 0 || y != nullptr), FHC_E_INVALID, "fhc_host_antitonic: bad arguments");
    if (m == 0) return FHC_OK;
    // stack of blocks (sum, number of points); the counts are kept as doubles (exact integers) so that the comparison
    // below needs no conversion per step; the buffers live as long as the calling thread
    static thread_local std::vector<double> sum_buf, cnt_buf;
    if ((int64_t)sum_buf.size() < m) {
        sum_buf.resize((size_t)m);
        cnt_buf.resize((size_t)m);
    }
    double *sum = sum_buf.data(), *cnt = cnt_buf.data();
    // The newest block lives in registers (ps, pc): a point that does not violate it costs one multiply and one compare
    // with no load that depends on the previous store -- most of a decaying spline table is such points.  The additions
    // happen in the order of the textbook loop (new point first, then the blocks below it, nearest first).
    int64_t top = -1;  // blocks below the newest one
    double ps = y[0], pc = 1.0;
    for (int64_t i = 1; i < m; ++i) {
        double s = y[i];
        if (!(ps < s * pc)) {  // non-increasing so far: the newest block is final for now
            ++top;
            sum[top] = ps;
            cnt[top] = pc;
            ps = s;
            pc = 1.0;
            continue;
        }
        // the previous block violates when its mean is below the new block's mean
        double c = 1.0 + pc;
        s += ps;
        while (top >= 0 && sum[top] * c < s * cnt[top]) {
            s += sum[top];
            c += cnt[top];
            --top;
        }
        ps = s;
        pc = c;
    }
    ++top;
    sum[top] = ps;
    cnt[top] = pc;
    int64_t i = 0;
    for (int64_t b = 0; b <= top; ++b) {
        if (cnt[b] == 1.0) {
            y[i++] = sum[b];
            continue;
        }
        const double mean = sum[b] / cnt[b];
        const int64_t n = (int64_t)cnt[b];
        for (int64_t k = 0; k < n; ++k) y[i++] = mean;
    }
    return FHC_OK;
}

extern "C" int fhc_spline_lut(const int64_t *splineX, const double *table, int64_t m, double xmin, double xmax,
                              int32_t res, double *lut, int64_t D, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(splineX && table && lut && m > 0 && D > 0 && res > 0, FHC_E_INVALID, "fhc_spline_lut: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    const int threads = 256;
    spline_lut_kernel<<<(unsigned int)((D + threads - 1) / threads), threads, 0, st>>>(
        reinterpret_cast<const long long *>(splineX), table, m, xmin, xmax, (unsigned int)res, lut, D);
    FHC_LAUNCH_CHECK("spline_lut_kernel");
    return FHC_OK;
}
