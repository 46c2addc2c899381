// Possible pairs per bin in restriction-fragment mode (-r 0) on the GPU, from prefix sums.
//
// Reference: generate_FragPairs, restriction-fragment branch (fithic/fithic.py:691-778).  Every pair (x < y) of one
// chromosome with L <= mid_y - mid_x <= U adds 1 to the bin's `[1]`, npairs = templen - d to `[7]` (d = in-range partners of
// x seen so far: the reference's quirk, :713-714) and float(dist / 1e6) * npairs to `[3]`.  The reference walks all pairs
// (O(n^2) per chromosome), fhc_host_frag_pairs_varsize (host_bins.cu) the pairs in range, one after the other so that the
// double sum `[3]` has the reference's bits -- fine for a window of a few Mb, hopeless for a genome-wide restriction map
// without -U (chr1 alone: 1e10 pairs).  Here a (fragment, bin) cell is O(log n):
//   * the mid points of a chromosome are sorted, so the partners of x that fall into bin b are one index range [ya, yb),
//     found by bisection (the bins are contiguous: bin(d) = the first b with d <= ub[b], clamped to the last -- the forward
//     tracker of :719-731);
//   * `[1]` += yb - ya;  `[7]` += sum_{y} (n + lo - y) in closed form (lo = first partner in range);
//   * `[3]`: sum_{y} (mid_y - mid_x) (n + lo - y) = (n + lo) * S1 - S2 - mid_x * `[7]`-term, with S1 = sum mid_y and
//     S2 = sum y * mid_y from prefix arrays -- exact integers (128 bit), divided by 1e6 once at the end.
// `[1]` and `[7]` equal the reference's bit for bit.  `[3]` is the exactly rounded value of the sum the reference
// accumulates term by term in double precision: it differs from the reference's bits by the rounding of that accumulation
// (relative 1e-13 on the bundled HindIII fragments; tests/test_gpu_kernels.py), far inside the 1e-6 the p-values allow.  The
// engine therefore keeps the host walk where it is cheap (FHC_VARSIZE_PAIRS, engine.py) and comes here beyond.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace fhc {

using u64 = unsigned long long;
using u128 = unsigned __int128;

struct FragPairsArgs {
    const long long *mids;     // all chromosomes back to back, sorted inside a chromosome
    const long long *chr_off;  // nchr + 1
    const u64 *p1;             // n_total + 1: prefix sums of mid
    const u64 *p2_lo, *p2_hi;  // n_total + 1: prefix sums of (index inside the chromosome) * mid, 128 bit
    const long long *bin_ub;   // nbins
    u64 *acc;                  // nbins x 4: [1], [7], [3] low word, [3] high word
    u64 *totals;               // [0] pairs in range, [1] largest distance in range
    long long n_total, L, U;
    int nchr, nbins;
};

// first index in [a, b) with f[i] >= v
__host__ __device__ __forceinline__ long long lower_bound_dev(const long long *__restrict__ f, long long a, long long b, long long v) {
    while (a < b) {
        const long long m = (a + b) >> 1;
        if (f[m] < v) a = m + 1; else b = m;
    }
    return a;
}

// first index in [a, b) with f[i] > v
__host__ __device__ __forceinline__ long long upper_bound_dev(const long long *__restrict__ f, long long a, long long b, long long v) {
    while (a < b) {
        const long long m = (a + b) >> 1;
        if (f[m] <= v) a = m + 1; else b = m;
    }
    return a;
}

// The partners [ya, yb) of fragment xi (mid point mx, first partner in range lo) of a chromosome with n fragments whose
// prefix arrays start at `off`: how many, the sum of npairs = n - (y - lo), and the sum of (mid_y - mx) * npairs.
__host__ __device__ __forceinline__ void frag_cell(const FragPairsArgs &a, long long off, long long n, long long lo,
                                                   long long ya, long long yb, long long mx, u64 *cnt_out, u64 *s7_out,
                                                   u128 *t_out) {
    const u64 cnt = (u64)(yb - ya);
    const u64 ja = (u64)(ya - lo);
    const u64 s7 = cnt * (u64)n - ((2 * ja + cnt - 1) * cnt) / 2;  // sum over j = ja .. ja + cnt - 1 of (n - j)
    const u64 s1 = a.p1[off + yb] - a.p1[off + ya];
    const u128 s2 = (((u128)a.p2_hi[off + yb] << 64) | a.p2_lo[off + yb]) - (((u128)a.p2_hi[off + ya] << 64) | a.p2_lo[off + ya]);
    *cnt_out = cnt;
    *s7_out = s7;
    *t_out = (u128)s1 * (u128)(u64)(n + lo) - s2 - (u128)(u64)mx * (u128)s7;
}

// the end of bin b's partner range for a fragment whose range so far ends at ya (window end hi)
__host__ __device__ __forceinline__ long long frag_bin_end(const FragPairsArgs &a, const long long *f, int b, long long ya,
                                                           long long hi, long long mx) {
    if (ya >= hi) return ya;
    if (b == a.nbins - 1) return hi;  // the tracker stops at the last bin
    const long long ub = a.bin_ub[b];
    return f[ya] - mx <= ub ? upper_bound_dev(f, ya + 1, hi, mx + ub) : ya;
}

__device__ __forceinline__ void add128(u64 *lo_hi, u64 lo, u64 hi) {
    if (lo) {
        const u64 old = atomicAdd(lo_hi, lo);
        hi += (old + lo < old) ? 1ull : 0ull;  // this addition wrapped the low word: carry (exact whatever the order)
    }
    if (hi) atomicAdd(lo_hi + 1, hi);
}

// One lane per fragment x, the 32 lanes of a warp walk the bins together: per bin every lane has its cell, the warp adds the
// 32 cells up and lane 0 adds the sum to the bin (integers: the order does not matter).
__global__ void __launch_bounds__(256) frag_pairs_varsize_kernel(const FragPairsArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u64 inrange = 0, maxdist = 0;
    for (long long base = warp0 * 32; base < a.n_total; base += warps * 32) {
        const long long gx = base + lane;
        const bool live = gx < a.n_total;
        long long off = 0, n = 0, xi = 0, mx = 0, lo = 0, hi = 0;
        if (live) {
            int c0 = 0, c1 = a.nchr;  // chromosome of gx: the last c with chr_off[c] <= gx
            while (c1 - c0 > 1) {
                const int m = (c0 + c1) >> 1;
                if (a.chr_off[m] <= gx) c0 = m; else c1 = m;
            }
            off = a.chr_off[c0];
            n = a.chr_off[c0 + 1] - off;
            xi = gx - off;
            const long long *f = a.mids + off;
            mx = f[xi];
            lo = a.L < 0 ? xi + 1 : lower_bound_dev(f, xi + 1, n, mx + a.L);   // in_range_check: -1 = unbounded
            hi = a.U < 0 ? n : upper_bound_dev(f, lo, n, mx + a.U);
            if (hi > lo) {
                inrange += (u64)(hi - lo);
                const u64 d = (u64)(f[hi - 1] - mx);
                maxdist = d > maxdist ? d : maxdist;
            }
        }
        if (a.nbins == 0) continue;
        const long long *f = a.mids + off;
        long long ya = lo;
        for (int b = 0; b < a.nbins; ++b) {
            if (__all_sync(0xffffffffu, ya >= hi)) break;  // every lane's window is used up
            const long long yb = frag_bin_end(a, f, b, ya, hi, mx);
            u64 cnt = (u64)(yb - ya);
            if (!__any_sync(0xffffffffu, cnt != 0)) continue;
            u64 s7 = 0, t_lo = 0, t_hi = 0;
            if (cnt) {
                u128 t;
                frag_cell(a, off, n, lo, ya, yb, mx, &cnt, &s7, &t);
                t_lo = (u64)t;
                t_hi = (u64)(t >> 64);
            }
            // 128-bit warp sum: low words with their carries
            const u64 c_sum = warp_sum(cnt), s7_sum = warp_sum(s7);
            u64 lo_sum = t_lo, hi_sum = t_hi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const u64 ol = __shfl_xor_sync(0xffffffffu, lo_sum, o);
                const u64 oh = __shfl_xor_sync(0xffffffffu, hi_sum, o);
                const u64 nl = lo_sum + ol;
                hi_sum += oh + (nl < lo_sum ? 1ull : 0ull);
                lo_sum = nl;
            }
            if (lane == 0) {
                u64 *cell = a.acc + 4 * (size_t)b;
                atomicAdd(cell, c_sum);
                atomicAdd(cell + 1, s7_sum);
                add128(cell + 2, lo_sum, hi_sum);
            }
            ya = yb;
        }
    }
    inrange = warp_sum(inrange);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 t = __shfl_xor_sync(0xffffffffu, maxdist, o);
        maxdist = t > maxdist ? t : maxdist;
    }
    if (lane == 0) {
        if (inrange) atomicAdd(a.totals, inrange);
        if (maxdist) atomicMax(a.totals + 1, maxdist);
    }
}

}  // namespace fhc

namespace {
struct FragPrefix {
    std::vector<long long> off;
    std::vector<fhc::u64> p1, p2_lo, p2_hi;
    int64_t n_total = 0, inter2 = 0, base = 0;
};

// prefix sums on the host (one pass over the fragments); the argument checks ride along
int frag_prefix(const char *who, const int64_t *mids, const int64_t *chr_off, int32_t nchr, const int64_t *bin_lb,
                const int64_t *bin_ub, int32_t nbins, const int64_t *bin_pairs1, const int64_t *bin_pairs7,
                const double *bin_sumdist, const int64_t *totals, FragPrefix *out) {
    using namespace fhc;
    FHC_REQUIRE(nchr >= 0 && nbins >= 0 && totals != nullptr, FHC_E_INVALID, "%s: bad nchr / nbins / totals", who);
    FHC_REQUIRE(nchr == 0 || (mids && chr_off), FHC_E_INVALID, "%s: null fragment arrays", who);
    FHC_REQUIRE(nbins == 0 || (bin_lb && bin_ub && bin_pairs1 && bin_pairs7 && bin_sumdist), FHC_E_INVALID, "%s: null bin arrays", who);
    for (int b = 1; b < nbins; ++b)
        FHC_REQUIRE(bin_lb[b] == bin_ub[b - 1] + 1, FHC_E_INVALID, "%s: bins are not contiguous at bin %d", who, b);
    out->base = nchr > 0 ? chr_off[0] : 0;
    const int64_t n_total = out->n_total = nchr > 0 ? chr_off[nchr] - out->base : 0;
    out->off.assign((size_t)nchr + 1, 0);
    out->p1.assign((size_t)n_total + 1, 0);
    out->p2_lo.assign((size_t)n_total + 1, 0);
    out->p2_hi.assign((size_t)n_total + 1, 0);
    u64 s1 = 0;
    u128 s2 = 0;
    for (int c = 0; c < nchr; ++c) {
        const int64_t *f = mids + chr_off[c];
        const int64_t n = chr_off[c + 1] - chr_off[c];
        FHC_REQUIRE(n >= 0, FHC_E_INVALID, "%s: chr_off decreases at chromosome %d", who, c);
        out->off[(size_t)c] = chr_off[c] - out->base;
        out->inter2 += (n_total - n) * n;  // :701
        for (int64_t i = 0; i < n; ++i) {
            FHC_REQUIRE(f[i] >= 0 && (i == 0 || f[i] >= f[i - 1]), FHC_E_INVALID,
                        "%s: mid points of chromosome %d are not sorted (or negative)", who, c);
            const size_t g = (size_t)(chr_off[c] - out->base + i);
            out->p1[g] = s1;
            out->p2_lo[g] = (u64)s2;
            out->p2_hi[g] = (u64)(s2 >> 64);
            s1 += (u64)f[i];
            s2 += (u128)(u64)i * (u128)(u64)f[i];
        }
    }
    out->off[(size_t)nchr] = n_total;
    out->p1[(size_t)n_total] = s1;
    out->p2_lo[(size_t)n_total] = (u64)s2;
    out->p2_hi[(size_t)n_total] = (u64)(s2 >> 64);
    return FHC_OK;
}

// acc: nbins x 4 words + [pairs in range, largest distance]
void frag_finish(const std::vector<fhc::u64> &acc, int32_t nbins, const FragPrefix &pre, int64_t *bin_pairs1,
                 int64_t *bin_pairs7, double *bin_sumdist, int64_t *totals) {
    using namespace fhc;
    const size_t w_acc = 4 * (size_t)nbins;
    for (int b = 0; b < nbins; ++b) {
        bin_pairs1[b] += (int64_t)acc[4 * (size_t)b];
        bin_pairs7[b] += (int64_t)acc[4 * (size_t)b + 1];
        const u128 t = ((u128)acc[4 * (size_t)b + 3] << 64) | acc[4 * (size_t)b + 2];
        bin_sumdist[b] += (double)t / 1000000.0;  // the conversion rounds the exact integer once
    }
    totals[0] = (int64_t)acc[w_acc];                  // possibleIntraInRangeCount
    totals[1] = nbins > 0 ? (int64_t)acc[w_acc] : 0;  // possibleIntraAllCount (:736: counted only when bins exist)
    totals[2] = pre.inter2;
    totals[3] = pre.n_total;
    totals[4] = (int64_t)acc[w_acc + 1];              // maxPossibleGenomicDist
}
}  // namespace

// Same arguments and results as fhc_host_frag_pairs_varsize (host arrays in, host arrays out; bin_pairs1 / bin_pairs7 carry
// the pass >= 2 outlier decrements on entry) plus the stream the copies and the kernel run on; synchronises that stream.
extern "C" int fhc_frag_pairs_varsize(const int64_t *mids, const int64_t *chr_off, int32_t nchr, int64_t L, int64_t U,
                                      const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins, int64_t *bin_pairs1,
                                      int64_t *bin_pairs7, double *bin_sumdist, int64_t *totals, void *stream) {
    using namespace fhc;
    FragPrefix pre;
    int rc = frag_prefix("fhc_frag_pairs_varsize", mids, chr_off, nchr, bin_lb, bin_ub, nbins, bin_pairs1, bin_pairs7,
                         bin_sumdist, totals, &pre);
    if (rc != FHC_OK) return rc;
    const int64_t n_total = pre.n_total;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    // one device block: [mids | chr_off | p1 | p2_lo | p2_hi | bin_ub | acc | totals]
    const size_t w_mids = (size_t)n_total, w_off = (size_t)nchr + 1, w_p = (size_t)n_total + 1, w_ub = (size_t)nbins;
    const size_t w_acc = 4 * (size_t)nbins, words = w_mids + w_off + 3 * w_p + w_ub + w_acc + 2;
    u64 *dev = nullptr;
    FHC_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dev), words * sizeof(u64), st));
    u64 *d_mids = dev, *d_off = d_mids + w_mids, *d_p1 = d_off + w_off, *d_p2l = d_p1 + w_p, *d_p2h = d_p2l + w_p;
    u64 *d_ub = d_p2h + w_p, *d_acc = d_ub + w_ub, *d_tot = d_acc + w_acc;
    std::vector<u64> acc(w_acc + 2, 0);
    do {
#define FHC_TRY(expr)                                                                                      \
    if ((expr) != cudaSuccess) {                                                                           \
        fhc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(cudaGetLastError()), __FILE__, __LINE__); \
        rc = FHC_E_CUDA;                                                                                   \
        break;                                                                                             \
    }
        if (n_total) FHC_TRY(cudaMemcpyAsync(d_mids, mids + pre.base, w_mids * 8, cudaMemcpyHostToDevice, st));
        FHC_TRY(cudaMemcpyAsync(d_off, pre.off.data(), w_off * 8, cudaMemcpyHostToDevice, st));
        FHC_TRY(cudaMemcpyAsync(d_p1, pre.p1.data(), w_p * 8, cudaMemcpyHostToDevice, st));
        FHC_TRY(cudaMemcpyAsync(d_p2l, pre.p2_lo.data(), w_p * 8, cudaMemcpyHostToDevice, st));
        FHC_TRY(cudaMemcpyAsync(d_p2h, pre.p2_hi.data(), w_p * 8, cudaMemcpyHostToDevice, st));
        if (nbins) FHC_TRY(cudaMemcpyAsync(d_ub, bin_ub, w_ub * 8, cudaMemcpyHostToDevice, st));
        FHC_TRY(cudaMemsetAsync(d_acc, 0, (w_acc + 2) * 8, st));
        if (n_total > 0) {
            FragPairsArgs a;
            a.mids = reinterpret_cast<const long long *>(d_mids);
            a.chr_off = reinterpret_cast<const long long *>(d_off);
            a.p1 = d_p1;
            a.p2_lo = d_p2l;
            a.p2_hi = d_p2h;
            a.bin_ub = reinterpret_cast<const long long *>(d_ub);
            a.acc = d_acc;
            a.totals = d_tot;
            a.n_total = n_total;
            a.L = L;
            a.U = U;
            a.nchr = nchr;
            a.nbins = nbins;
            long long blocks = (n_total + 255) / 256;
            if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
            frag_pairs_varsize_kernel<<<(unsigned int)blocks, 256, 0, st>>>(a);
            if (cudaGetLastError() != cudaSuccess) {
                fhc::set_error("launch of frag_pairs_varsize_kernel failed");
                rc = FHC_E_CUDA;
                break;
            }
            fhc::count_launch();
            if (fhc::g_profile_on) fhc::profile_mark("frag_pairs_varsize_kernel", st);
        }
        FHC_TRY(cudaMemcpyAsync(acc.data(), d_acc, (w_acc + 2) * 8, cudaMemcpyDeviceToHost, st));
        FHC_TRY(cudaStreamSynchronize(st));
#undef FHC_TRY
    } while (0);
    cudaFreeAsync(dev, st);
    if (rc != FHC_OK) return rc;
    frag_finish(acc, nbins, pre, bin_pairs1, bin_pairs7, bin_sumdist, totals);
    return FHC_OK;
}

// The kernel's cells, one after the other on the host (what the CPU tests check against the pair-by-pair walk of
// fhc_host_frag_pairs_varsize).
extern "C" int fhc_host_frag_pairs_varsize_prefix(const int64_t *mids, const int64_t *chr_off, int32_t nchr, int64_t L, int64_t U,
                                                  const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins,
                                                  int64_t *bin_pairs1, int64_t *bin_pairs7, double *bin_sumdist,
                                                  int64_t *totals) {
    using namespace fhc;
    FragPrefix pre;
    const int rc = frag_prefix("fhc_host_frag_pairs_varsize_prefix", mids, chr_off, nchr, bin_lb, bin_ub, nbins, bin_pairs1,
                               bin_pairs7, bin_sumdist, totals, &pre);
    if (rc != FHC_OK) return rc;
    FragPairsArgs a;
    a.mids = reinterpret_cast<const long long *>(mids + pre.base);
    a.chr_off = pre.off.data();
    a.p1 = pre.p1.data();
    a.p2_lo = pre.p2_lo.data();
    a.p2_hi = pre.p2_hi.data();
    a.bin_ub = reinterpret_cast<const long long *>(bin_ub);
    a.acc = nullptr;
    a.totals = nullptr;
    a.n_total = pre.n_total;
    a.L = L;
    a.U = U;
    a.nchr = nchr;
    a.nbins = nbins;
    const size_t w_acc = 4 * (size_t)nbins;
    std::vector<u64> acc(w_acc + 2, 0);
    std::vector<u128> sums((size_t)nbins, 0);
    for (int c = 0; c < nchr; ++c) {
        const long long off = pre.off[(size_t)c], n = pre.off[(size_t)c + 1] - off;
        const long long *f = a.mids + off;
        for (long long xi = 0; xi < n; ++xi) {
            const long long mx = f[xi];
            const long long lo = L < 0 ? xi + 1 : lower_bound_dev(f, xi + 1, n, mx + L);
            const long long hi = U < 0 ? n : upper_bound_dev(f, lo, n, mx + U);
            if (hi <= lo) continue;
            acc[w_acc] += (u64)(hi - lo);
            const u64 d = (u64)(f[hi - 1] - mx);
            acc[w_acc + 1] = d > acc[w_acc + 1] ? d : acc[w_acc + 1];
            long long ya = lo;
            for (int b = 0; b < nbins && ya < hi; ++b) {
                const long long yb = frag_bin_end(a, f, b, ya, hi, mx);
                if (yb == ya) continue;
                u64 cnt, s7;
                u128 t;
                frag_cell(a, off, n, lo, ya, yb, mx, &cnt, &s7, &t);
                acc[4 * (size_t)b] += cnt;
                acc[4 * (size_t)b + 1] += s7;
                sums[(size_t)b] += t;
                ya = yb;
            }
        }
    }
    for (int b = 0; b < nbins; ++b) {
        acc[4 * (size_t)b + 2] = (u64)sums[(size_t)b];
        acc[4 * (size_t)b + 3] = (u64)(sums[(size_t)b] >> 64);
    }
    frag_finish(acc, nbins, pre, bin_pairs1, bin_pairs7, bin_sumdist, totals);
    return FHC_OK;
}
